#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1
timeout 300 python -m pytest tests -q -m gpu -k "odd" > gpurun_out/j_tests.log 2>&1; tail -3 gpurun_out/j_tests.log
for spec in "pspnet 713 16" "psanet 473 16"; do
  set -- $spec
  timeout 400 python bench.py --model $1 --size $2 --batch $3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_$1.json 2> gpurun_out/j_bench_$1.err; echo "$1 rc=$?"
  grep "^{" gpurun_out/j_bench_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), round(d['ms_per_step'],2), d['config']['step_launch'], 'conv TF/s', round(d['roofline']['achieved'],1), 'eager', round(d['config']['eager_ms_per_step'],2))" 2>/dev/null || tail -3 gpurun_out/j_bench_$1.err
done
