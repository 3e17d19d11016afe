#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bisenet.py -q -m gpu -p no:cacheprovider -k "eval_forward" 2>&1 | tail -12 | cut -c1-220
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3 | cut -c1-300
