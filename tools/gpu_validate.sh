#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z_build.log 2>&1
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/z_bench.json | head -c 400; echo
