#!/bin/bash
# round-2 GPU call C (1 GPU): fixed tests at full size, bench eager + graph
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -s -k "fullsize or fcn or full_size or taps or psp or psa" > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
grep -n "full-size\|head [0-9]\|worst\|outside\|cosine\|running stat\|regime\|passed\|failed\|Error\|^parameter\|^heads\|^arms\|^ffm\|^refines\|^spatial\|^global\|^context_path.conv1\|^context_path.layer[1-4].0.conv1" gpurun_out/c_tests.log | head -80
timeout 300 python bench.py --steps 10 --warmup 3 --graph --no-cpu-baseline > gpurun_out/c_bench_graph.json 2> gpurun_out/c_bench_graph.err; echo "bench rc=$?" >> gpurun_out/c_bench_graph.err
tail -3 gpurun_out/c_bench_graph.err; head -c 900 gpurun_out/c_bench_graph.json; echo
