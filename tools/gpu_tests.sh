#!/bin/bash
# run the GPU test files in separate processes (a sticky CUDA error must not poison later files)
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=${CUDA_LAUNCH_BLOCKING:-1}
for f in test_gpu_kernels test_gpu_conv test_gpu_bisenet; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider "$@" > gpurun_out/$f.log 2>&1
  echo "== $f rc=$? =="; tail -n 40 gpurun_out/$f.log | grep -E "passed|failed|FAILED|Error|error" | head -60
done
