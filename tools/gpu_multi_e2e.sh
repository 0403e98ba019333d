#!/bin/bash
# e2e limiter at N GPUs: the default bench line, then the e2e part again with write-combined pinned input and with
# 768 MiB waves (B200SDR_PINNED_WC / B200SDR_WAVE_MB, measurement knobs of api.cu)
NG=${NG:-8}
bash tools/gpu_multi.sh
for V in "B200SDR_PINNED_WC=1" "B200SDR_WAVE_MB=768" "B200SDR_WAVE_MB=48"; do
  env $V timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --steps 3 --warmup 3 --captures-per-gpu 64 --e2e-captures 64 --parity-captures 2 > gpurun_out/bench_e2e_var.txt 2> gpurun_out/bench_e2e_var.err
  python - "$V" <<'PY' | tee -a gpurun_out/e2e_variants_n$NG.txt
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/bench_e2e_var.txt').read().strip().splitlines() if l.startswith('{')][-1])
    e=d['e2e']; print(sys.argv[1], 'N', d['n_gpus'], 'e2e MS/s', round(e['value']), 'GB/s agg', round(e['h2d_GBps_aggregate'],1), 'per rank', e['h2d_GBps_per_rank'], '| ceiling agg', round(e['h2d_ceiling']['GBps_aggregate'],1), e['h2d_ceiling']['GBps_per_rank'], 'frac', round(e['frac_of_h2d_ceiling'],3))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open('gpurun_out/bench_e2e_var.err').read()[-800:])
PY
done
