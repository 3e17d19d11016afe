#!/bin/bash
# 1 GPU: round-2 evidence — default bench, launch list of one eager step (time + DRAM bytes), ncu --set full of the hot kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/k_build.log 2>&1
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/k_bench.json | head -c 500; echo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/k_launches.csv python tools/profile_step.py > gpurun_out/k_profile_step.log 2>&1; echo "ncu rc=$?"
python tools/agg_launches.py gpurun_out/k_launches.csv | head -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'igemm_v2|wgrad_taps|ohem_ptarget_up|ohem_grad_up' -s 6 -c 6 -f -o gpurun_out/r02_hot2 python tools/prof_one.py > gpurun_out/k_ncu.log 2>&1; echo "ncu full rc=$?"
