#!/bin/bash
# Run on the GPU box (through gpurun): each test file in its own process with a timeout so that a
# trapped / hung kernel cannot take the rest of the run down with it.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in "$@"; do
  name=$(basename "$t" .py)
  echo "=== $t"
  timeout 600 python -m pytest "$t" -q -s -m gpu -x --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "exit $?"; tail -5 "gpurun_out/$name.log"
done
