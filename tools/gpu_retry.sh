#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): tools/gpu_retry.sh <timeout> <gpus> '<command>'
T=$1; G=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  if [ "$G" = 1 ]; then gpurun --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; else gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; fi
  if ! grep -q "status=transient" gpurun_out/.retry.log; then cat gpurun_out/.retry.log; exit 0; fi
  sleep 90
done
cat gpurun_out/.retry.log; exit 3
