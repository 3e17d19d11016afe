#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x -k "not graphed" > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log; tail -15 gpurun_out/h_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/h_bench.json | head -c 600; echo
TSB_DEBUG_SET="10=0" timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_oldohem.json 2> gpurun_out/h_bench_oldohem.err; grep "^{" gpurun_out/h_bench_oldohem.json | head -c 300; echo
timeout 400 python bench.py --eval --steps 20 > gpurun_out/h_bench_eval.json 2> gpurun_out/h_bench_eval.err; echo "eval rc=$?"; cat gpurun_out/h_bench_eval.json | head -c 1500; echo; tail -3 gpurun_out/h_bench_eval.err
