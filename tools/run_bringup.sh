#!/bin/bash
# run each bring-up step in its own process under a timeout, so a hung kernel cannot eat the GPU lease
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
for step in "$@"; do
  echo "=== $step"
  timeout -s KILL ${STEP_TIMEOUT:-120} python tools/gpu_bringup.py $step 2>&1 | tail -60
  echo "=== $step exit ${PIPESTATUS[0]}"
done
