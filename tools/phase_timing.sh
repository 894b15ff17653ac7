#!/bin/bash
# dev tool: build a phase-timing variant of the library and print cycles per phase per work item
set -e
cd "$(dirname "$0")/../deep_cine_cardiac_mri_b200/csrc"
mkdir -p build_dbg
for f in b2s_abi b2s_fused b2s_generic b2s_pointwise b2s_normal; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DB2S_PHASE_TIMING -c $f.cu -o build_dbg/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_dbg/libb2sense_dbg.so build_dbg/*.o -lcudart
