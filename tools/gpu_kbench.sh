#!/bin/bash
# On the GPU box (through gpurun): kernel-only ranking of libhbk builds on a one-wave C2-shaped launch (tools/kbench.py).
#   usage: bash tools/gpu_kbench.sh <tag> [variant ...]     variant = name of variants/libhbk_<name>.so, "default" = the
#   in-tree library. Output: gpurun_out/<tag>_kbench.jsonl (one line per variant and workload).
# Env: KB_ARGS (extra kbench arguments), KB_NCU=<variant> (ncu --set full source capture of that variant's hot kernel),
#      KB_TESTS=1 (run the -m gpu tests on the default library first)
T=${1:-r02}; shift
V=${@:-default}
O=gpurun_out
mkdir -p $O
if [ -n "$KB_TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -q -x > $O/${T}_gpu_tests.log 2>&1; tail -3 $O/${T}_gpu_tests.log
fi
for v in $V; do
  lib=variants/libhbk_$v.so; [ "$v" = default ] && lib=hydrobricks_b200/libhbk.so
  HBK_LIB=$PWD/$lib timeout 120 python tools/kbench.py --tag $v $KB_ARGS >> $O/${T}_kbench.jsonl 2>> $O/${T}_kbench.err
  tail -1 $O/${T}_kbench.jsonl | cut -c1-330
done
if [ -n "$KB_NCU" ]; then
  lib=variants/libhbk_$KB_NCU.so; [ "$KB_NCU" = default ] && lib=hydrobricks_b200/libhbk.so
  HBK_LIB=$PWD/$lib timeout 300 ncu --set full --clock-control none --import-source on -k regex:hbkRun -c 1 -f \
      -o $O/${T}_main python tools/kbench.py --reps 0 --timesteps 730 $KB_ARGS > $O/${T}_ncu.log 2>&1
  ncu -i $O/${T}_main.ncu-rep --page source --csv > $O/${T}_main_source.csv 2>/dev/null
  ncu -i $O/${T}_main.ncu-rep --page raw --csv > $O/${T}_main_raw.csv 2>/dev/null
  ls -la $O/${T}_main*
fi
