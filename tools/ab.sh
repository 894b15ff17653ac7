#!/bin/bash
# dev tool: A/B of kernel variants on one GPU box.  usage: tools/ab.sh tag "ENV1=a ENV2=b" "ENV1=c" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  echo "=== variant $i: $envs" >> gpurun_out/${tag}_ab.log
  env $envs python tools/quick_bench.py >> gpurun_out/${tag}_ab.log 2>&1
  env $envs python tools/pipe_time.py >> gpurun_out/${tag}_ab.log 2>&1
  i=$((i+1))
done
cat gpurun_out/${tag}_ab.log
