#!/bin/bash
# On an N-GPU box (gpurun --gpus N): the C2 bench at N ranks, weak (100 000 sets per GPU) and strong (100 000 sets in total).
#   usage: bash tools/gpu_scaling.sh <tag> <N>        -> gpurun_out/<tag>_bench_{weak,strong}_<N>gpu.json
T=${1:-r02}; N=${2:-2}
O=gpurun_out
mkdir -p $O
for s in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling $s > $O/${T}_bench_${s}_${N}gpu.json 2> $O/${T}_bench_${s}_${N}gpu.err
  python -c "
import json
try:
    d=json.loads(open('$O/${T}_bench_${s}_${N}gpu.json').read().strip().splitlines()[-1]); print('$s', $N, '%.4g' % d['value'], '%.1f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'])
except Exception as e: print('$s', $N, 'FAILED', e)"
done
