#!/bin/bash
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1
timeout 900 python -m pytest tests/test_gpu_bisenet.py -m gpu -q --timeout 400 -p no:cacheprovider -k "pspnet" 2>&1 | tail -3
unset CUDA_LAUNCH_BLOCKING
timeout 900 python bench.py --model pspnet --batch 16 --steps 3 --warmup 3 > gpurun_out/bench_pspnet.json 2> gpurun_out/bench_pspnet.err; echo "pspnet bench rc=$?"; tail -3 gpurun_out/bench_pspnet.err | cut -c1-300; cut -c1-1200 gpurun_out/bench_pspnet.json
