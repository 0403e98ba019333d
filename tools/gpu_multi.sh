#!/bin/bash
# N-GPU check of the torchrun contract (N from $NG): topology dump + bench.py + the parity / e2e parts of its line
NG=${NG:-2}
mkdir -p gpurun_out
{ nvidia-smi -L | wc -l; cat /sys/devices/system/node/online; nproc; nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name"; 
  for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo "$(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done; } > gpurun_out/topo_n$NG.txt 2>&1
EXTRA="${EXTRA:-}"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps ${STEPS:-3} --warmup 3 $EXTRA > gpurun_out/bench_n$NG.txt 2> gpurun_out/bench_n$NG.err
python - $NG <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_n{n}.txt').read().strip().splitlines() if l.startswith('{')][-1])
    print('N',d['n_gpus'],'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
    print('parity', d['parity']['bitwise'], d['parity']['captures_recomputed_alone'], d['parity']['equal_to_batch_result'], d['parity']['probe_digest_per_rank'], d['parity']['probe_digest_in_rank0_batch'])
    e=d['e2e']; print('e2e GB/s agg', round(e['h2d_GBps_aggregate'],1), 'per rank', e['h2d_GBps_per_rank'], 'ceiling agg', round(e['h2d_ceiling']['GBps_aggregate'],1), 'per rank', e['h2d_ceiling']['GBps_per_rank'], 'frac', round(e['frac_of_h2d_ceiling'],3))
    print('split', d.get('split_capture'))
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_n{n}.err').read()[-2000:])
PY
head -40 gpurun_out/topo_n$NG.txt
