#!/bin/bash
# Round-end measurement suite for one B200 box (run through gpurun): parity tests, smoke, benches, ncu evidence.
# usage: bash tools/gpu_suite.sh <tag>      -> files gpurun_out/<tag>_*
T=${1:-r01l}
O=gpurun_out
timeout 120 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 150 python bench.py --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 90 python bench.py --no-cpu-baseline --workload c4 --steps 2 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
timeout 60 python bench.py --no-cpu-baseline --solver crank_nicolson --sets 4096 --timesteps 3653 --steps 2 \
    > $O/${T}_bench_c2small_crank_nicolson.json 2> $O/${T}_bench_cn.err
timeout 40 python bench.py --sets 28416 --timesteps 730 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
timeout 90 ncu --set full --clock-control none --import-source on -k regex:hbkRunUnits -c 1 -f -o $O/${T}_main \
    python bench.py --sets 28416 --timesteps 730 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for f in $O/${T}_bench*.json; do python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
except Exception as e: print('$f', 'FAILED', e)"; done
