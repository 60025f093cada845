#!/bin/bash
# Build the staged compile-time variants of libhbk.so next to the default one (here, no GPU needed):
#   variants/libhbk_faithful.so (-DHBK_FAITHFUL=1)   variants/libhbk_opt7.so (-DHBK_OPT=7)
# The *.so files are git-ignored but travel with the gpurun snapshot; tools/gpu_variants.sh measures them on the box.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared"
nvcc $FLAGS -DHBK_FAITHFUL=1 -o variants/libhbk_faithful.so hydrobricks_b200/csrc/hbk_engine.cu &
nvcc $FLAGS -DHBK_OPT=7 -o variants/libhbk_opt7.so hydrobricks_b200/csrc/hbk_engine.cu &
wait
ls -la variants/
