#!/bin/bash
# 2-GPU box: whole GPU suite in one process (includes tests/test_gpu_ddp.py) + the 2-GPU bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rs > gpurun_out/test_gpu_all_n2.log 2>&1
echo "== all rc=$? =="; grep -E "passed|failed|FAILED|Error|SKIPPED" gpurun_out/test_gpu_all_n2.log | cut -c1-200 | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err | cut -c1-300; cut -c1-600 gpurun_out/bench_n2.json
