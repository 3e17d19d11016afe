#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_gpu_conv.log 2>&1
echo "== conv rc=$? =="; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_gpu_conv.log | head -20
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_conv.py > gpurun_out/test_gpu_all.log 2>&1
echo "== all rc=$? =="; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
TSB_DEBUG_SET="7=0" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2>> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_a", "bench_b"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --profile-from-start off -k regex:igemm --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_igemm.csv python tools/profile_step.py > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log
python tools/agg_launches.py gpurun_out/launches_igemm.csv 2>/dev/null | head -10
