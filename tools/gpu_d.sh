#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 900 python tools/diag_graph.py > gpurun_out/d_diag_graph.log 2>&1; cat gpurun_out/d_diag_graph.log
timeout 300 python -m pytest tests/test_gpu_fcn.py -q > gpurun_out/d_fcn.log 2>&1; tail -3 gpurun_out/d_fcn.log
