#!/bin/bash
# round-2 GPU call B: new wgrad taps kernel + 256-bit epilogue (tests, A/B micro-bench), fixed tests, graph diagnosis
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x > gpurun_out/b_conv_tests.log 2>&1; echo "conv tests rc=$?" >> gpurun_out/b_conv_tests.log
tail -3 gpurun_out/b_conv_tests.log
timeout 300 python tools/bench_conv.py 8=0 9=0 > gpurun_out/b_benchconv_old.log 2>&1
timeout 300 python tools/bench_conv.py 8=1 9=0 > gpurun_out/b_benchconv_taps.log 2>&1
timeout 300 python tools/bench_conv.py 8=1 9=1 > gpurun_out/b_benchconv_taps_wide.log 2>&1
paste -d'\n' gpurun_out/b_benchconv_old.log gpurun_out/b_benchconv_taps.log gpurun_out/b_benchconv_taps_wide.log
timeout 900 python -m pytest tests -q -m gpu -s -k "fullsize or fcn" > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
grep -n "full-size\|head [0-9]\|worst\|norm_err >\|cosine <\|running stat\|regime\|passed\|failed\|Error" gpurun_out/b_tests.log | head -40
timeout 900 python tools/diag_graph.py > gpurun_out/b_diag_graph.log 2>&1; cat gpurun_out/b_diag_graph.log
