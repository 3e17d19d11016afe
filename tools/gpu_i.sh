#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu -k "not graphed" > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log; tail -12 gpurun_out/i_tests.log
TSB_TEST_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_bisenet.py -q -k graphed > gpurun_out/i_graph_test.log 2>&1; tail -2 gpurun_out/i_graph_test.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; tail -2 gpurun_out/i_smoke.log
