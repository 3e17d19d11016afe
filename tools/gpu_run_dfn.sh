#!/bin/bash
# DFN parity tests + regression of the BiSeNet/PSPNet suite + a DFN bench line
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=${CUDA_LAUNCH_BLOCKING:-1}
for f in test_gpu_dfn test_gpu_bisenet; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f rc=$? =="; grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/$f.log | head -40
done
unset CUDA_LAUNCH_BLOCKING
timeout 600 python bench.py --model dfn --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dfn.json 2> gpurun_out/bench_dfn.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_dfn.err; cat gpurun_out/bench_dfn.json
