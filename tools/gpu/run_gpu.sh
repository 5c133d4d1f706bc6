#!/bin/bash
# usage: tools/gpu/run_gpu.sh <tag> <timeout_s> <gpus> <command...>: gpurun with retries while the pod is busy (rc 3 / transient)
tag=$1; to=$2; gpus=$3; shift 3
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/${tag}_stdout.log 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > gpurun_out/${tag}_stdout.log 2>&1
  fi
  rc=$?
  if grep -q "status=transient\|status=busy\|nothing was charged" gpurun_out/${tag}_stdout.log || [ $rc -eq 3 ]; then
    sleep 90; continue
  fi
  break
done
echo "run_gpu $tag finished rc=$rc after $i attempt(s)" >> gpurun_out/${tag}_stdout.log
