#!/bin/bash
# -DCW_TIMING build of the library (per-CTA / per-warp %globaltimer stamps) for the timeline probes: ab/lib_timing.so
set -e
cd "$(dirname "$0")/.."
mkdir -p ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC --use_fast_math -DCW_TIMING \
     -I include -I gym_craftingworld_b200/csrc -o ab/lib_timing.so gym_craftingworld_b200/csrc/cw_kernels.cu gym_craftingworld_b200/csrc/cw_host.cu
echo ab/lib_timing.so
