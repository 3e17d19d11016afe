#!/bin/bash
# round-2 GPU call A: full GPU test suite, graph experiment, bench with reference_gpu
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -s -k "not graphed" > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
TSB_TEST_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_bisenet.py -q -k graphed > gpurun_out/a_graph_test.log 2>&1; echo "rc=$?" >> gpurun_out/a_graph_test.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --graph --no-cpu-baseline > gpurun_out/a_bench_graph.json 2> gpurun_out/a_bench_graph.err; echo "bench rc=$?" >> gpurun_out/a_bench_graph.err
tail -5 gpurun_out/a_tests.log; tail -3 gpurun_out/a_graph_test.log; cat gpurun_out/a_bench.json | head -c 1500; echo; tail -3 gpurun_out/a_bench_graph.err; head -c 600 gpurun_out/a_bench_graph.json
