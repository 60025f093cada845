#!/bin/bash
# On the GPU box (through gpurun), after tools/build_variants.sh: parity tests + C2 bench for every variant of libhbk.so,
# each swapped in for the default library on the box's scratch copy.   usage: bash tools/gpu_variants.sh <tag> [variant...]
T=${1:-r02a}; shift
V=${@:-faithful opt7}
O=gpurun_out
cp hydrobricks_b200/libhbk.so variants/libhbk_default.so
for v in default $V; do
  cp variants/libhbk_$v.so hydrobricks_b200/libhbk.so
  timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -1
  timeout 150 python bench.py --no-cpu-baseline --steps 2 > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
  python -c "
import json
try:
    d=json.loads(open('$O/${T}_bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('$v', 'FAILED', e)"
done
cp variants/libhbk_default.so hydrobricks_b200/libhbk.so
