#!/bin/bash
# dev tool: phase cycle floors with loads and/or stores stubbed (fft2c kernel only)
set -e
cd "$(dirname "$0")/../deep_cine_cardiac_mri_b200/csrc"
for v in "" "-DB2S_NOLOAD" "-DB2S_NOSTORE" "-DB2S_NOLOAD -DB2S_NOSTORE"; do
  d=build_dbg_$(echo "$v" | tr -d ' -' ); mkdir -p $d
  for f in b2s_abi b2s_fused b2s_generic b2s_pointwise b2s_normal; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DB2S_PHASE_TIMING $v -c $f.cu -o $d/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/lib.so $d/*.o -lcudart
  echo "== variant [$v]"
  for nt in 256 512; do echo " NT=$nt"; B2S_NT=$nt B2S_LIB=$PWD/$d/lib.so python ../../tools/phase_probe.py 2>&1 | head -1; done
done
