#!/bin/bash
# N-GPU check of the torchrun contract (N from $NG)
NG=${NG:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; cat /sys/devices/system/node/online; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_n$NG.txt 2> gpurun_out/bench_n$NG.err
python - $NG <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_n{n}.txt').read().strip().splitlines() if l.startswith('{')][-1])
    print('N',d['n_gpus'],'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_n{n}.err').read()[-2000:])
PY
