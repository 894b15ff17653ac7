#!/bin/bash
# dev tool: cycles per phase for kernel variants (needs tools/phase_timing.sh run in the container first)
tag=$1; shift
mkdir -p gpurun_out
for envs in "$@"; do
  echo "=== $envs" >> gpurun_out/${tag}_phase.log
  env $envs B2S_LIB=$PWD/deep_cine_cardiac_mri_b200/csrc/build_dbg/libb2sense_dbg.so python tools/phase_probe.py >> gpurun_out/${tag}_phase.log 2>&1
done
cat gpurun_out/${tag}_phase.log
