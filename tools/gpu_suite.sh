#!/bin/bash
# Measurement suite for one B200 box (run through gpurun): parity tests, smoke, benches, ncu evidence.
# usage: bash tools/gpu_suite.sh <tag> [quick]      -> files gpurun_out/<tag>_*
T=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
if [ "$2" != quick ]; then
  # launch list of the bench command (ncu gpu__time_duration: share of the step per kernel)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  # DRAM traffic of the full-size C2 launch
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:hbkRunUnits -c 1 --csv --log-file $O/${T}_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
  # full counters + source page of one wave of the hot kernel
  KB_NCU=default bash tools/gpu_kbench.sh $T default
  for w in c3 c4 c5; do
    timeout 300 python bench.py --workload $w --steps 2 --warmup 1 > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  done
  for s in heun_explicit euler_explicit; do
    timeout 200 python bench.py --no-cpu-baseline --solver $s --steps 2 --warmup 1 > $O/${T}_bench_c2_$s.json 2> $O/${T}_bench_c2_$s.err
  done
  timeout 200 python bench.py --no-cpu-baseline --solver crank_nicolson --sets 4096 --timesteps 3653 --steps 2 --warmup 1 \
      > $O/${T}_bench_c2small_crank_nicolson.json 2> $O/${T}_bench_cn.err
fi
for f in $O/${T}_bench*.json; do python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.4g' % d['value'], '%.1f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('$f', 'FAILED', e)"; done
