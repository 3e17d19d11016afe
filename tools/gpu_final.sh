#!/bin/bash
# round-end consolidation: whole GPU suite in one process (the driver's form), default bench, launch list, ncu --set full
# of the pair kernel, secondary model lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
echo "== all rc=$? =="; grep -E "passed|failed|FAILED|Error|assert|SKIP|skipped" gpurun_out/test_gpu_all.log | head -20
timeout 600 python -m pytest tests/test_gpu_bisenet.py -q -m gpu -p no:cacheprovider -k graphed -rs > gpurun_out/test_graphed.log 2>&1; tail -30 gpurun_out/test_graphed.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log
ncu --profile-from-start off -k regex:"igemm_v2_kernel<128, 0, 1>" --launch-skip 4 --launch-count 3 --set full --import-source on --clock-control none -o gpurun_out/pair_full -f python tools/profile_step.py > gpurun_out/prof_pair.log 2>&1; tail -1 gpurun_out/prof_pair.log
for m in pspnet psanet; do timeout 600 python bench.py --model $m --batch 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "$m rc=$?"; tail -2 gpurun_out/bench_$m.err | cut -c1-300; done
timeout 600 python bench.py --model dfn --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dfn.json 2> gpurun_out/bench_dfn.err; echo "dfn rc=$?"; tail -2 gpurun_out/bench_dfn.err | cut -c1-300
python - <<'PY'
import json
for f in ("bench_default", "bench_pspnet", "bench_psanet", "bench_dfn"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
python tools/agg_launches.py gpurun_out/launches.csv 2>/dev/null | head -24
