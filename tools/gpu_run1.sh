#!/bin/bash
# tests + smoke + first bench line
mkdir -p gpurun_out
bash tools/gpu_tests.sh
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
