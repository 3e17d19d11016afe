#!/bin/bash
mkdir -p gpurun_out
# the driver's form: the whole GPU suite in ONE process
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
echo "== all rc=$? =="; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -30
