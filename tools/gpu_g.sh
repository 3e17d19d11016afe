#!/bin/bash
# 1 GPU: ncu --set full of the hot kernels (second pass of the loop: warm)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'igemm_v2|wgrad_taps|ohem_ptarget_up|ohem_grad_up' -s 6 -c 6 -f -o gpurun_out/r02_hot python tools/prof_one.py > gpurun_out/g_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/g_ncu.log
ls -la gpurun_out/r02_hot.ncu-rep
