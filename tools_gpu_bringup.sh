#!/bin/bash
# First-contact run on the B200 box: each test file in its own process so a faulting kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_kernels.py tests/test_gpu_path.py; do
  echo "=== $f" >> gpurun_out/bringup.log
  timeout 600 python -m pytest $f -m gpu -q --no-header --tb=short -p no:cacheprovider "$@" >> gpurun_out/bringup.log 2>&1
  echo "exit $?" >> gpurun_out/bringup.log
done
grep -vE '^E    (\+|   )' gpurun_out/bringup.log | tail -n 120
