#!/bin/bash
# 1 GPU: default bench (graph), launch list of one eager step with DRAM bytes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/f_bench.json | head -c 700; echo; tail -2 gpurun_out/f_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python tools/profile_step.py > gpurun_out/f_profile_step.log 2>&1; echo "ncu rc=$?"
python tools/agg_launches.py gpurun_out/f_launches.csv | head -40
