import importlib, sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); 
pkg = importlib.import_module("stm32f7-rtlsdr_b200")
L = 48_000_000
a = pkg.B200Sdr(chains=pkg.CHAIN_WBFM, fir_engine=pkg.FIR_ENGINE_TENSOR)
b = pkg.B200Sdr(chains=pkg.CHAIN_WBFM, fir_engine=pkg.FIR_ENGINE_FP32)
iq = torch.empty(4 * L, dtype=torch.uint8, device="cuda")
na = pkg.wbfm_audio_len(L)
x = torch.zeros(4 * na, dtype=torch.float32, device="cuda"); y = torch.zeros(4 * na, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for c in range(4): a.synth_fill_dev(iq.data_ptr() + c * L, 1, L, pkg.SYNTH_WBFM if c % 2 else pkg.SYNTH_MULTITONE, first_capture=c)
a.sync()
a.batch_wbfm_dev(iq.data_ptr(), 4, L, x.data_ptr()); a.sync()
b.batch_wbfm_dev(iq.data_ptr(), 4, L, y.data_ptr()); b.sync()
torch.cuda.synchronize()
print("max abs diff", float((x - y).abs().max()), "max abs", float(x.abs().max()), float(y.abs().max()), "per capture", [(float((x[c*na:(c+1)*na]-y[c*na:(c+1)*na]).abs().max())) for c in range(4)])
