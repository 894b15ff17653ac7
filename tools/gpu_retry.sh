#!/bin/bash
# dev tool: keep asking for a GPU box until the call is accepted (exit code 3 = no slot right now)
# usage: tools/gpu_retry.sh <timeout> '<command>'
to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
