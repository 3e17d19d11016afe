#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
echo "== all rc=$? =="; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench_a",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log
python tools/agg_launches.py gpurun_out/launches.csv 2>/dev/null | head -36
