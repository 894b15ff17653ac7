#!/bin/bash
# dev tool: build a phase-timing variant of the library (build_dbg/libb2sense_dbg.so); run tools/phase_probe.py with
# B2S_LIB pointing at it to print cycles per phase per work item
set -e
cd "$(dirname "$0")/../deep_cine_cardiac_mri_b200/csrc"
mkdir -p build_dbg
for f in b2s_abi b2s_fused b2s_strip b2s_generic b2s_pointwise b2s_normal b2s_metrics; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DB2S_PHASE_TIMING $B2S_EXTRA -c $f.cu -o build_dbg/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_dbg/libb2sense_dbg.so build_dbg/*.o -lcudart
