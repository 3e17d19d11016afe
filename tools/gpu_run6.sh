#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off -k regex:"bn_(bwd_reduce|bwd_apply|apply)_kernel" --launch-skip 6 --launch-count 9 --set full --import-source on --clock-control none -o gpurun_out/bn_full -f python tools/profile_step.py > gpurun_out/prof_bn.log 2>&1; tail -1 gpurun_out/prof_bn.log
