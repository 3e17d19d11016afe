#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bisenet.py tests/test_gpu_kernels.py tests/test_gpu_dfn.py -q -m gpu -p no:cacheprovider -k "graphed or prefetcher or mirror or dfn_r101" > gpurun_out/test_new.log 2>&1
echo "== new rc=$? =="; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_new.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench_a",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "eager", round(d["config"]["eager_ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "launches", d["gpu_launches"], d["config"]["step_launch"], d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
for m in pspnet psanet; do timeout 600 python bench.py --model $m --batch 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "$m rc=$?"; tail -2 gpurun_out/bench_$m.err | cut -c1-300; done
timeout 600 python bench.py --model dfn --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dfn.json 2> gpurun_out/bench_dfn.err; echo "dfn rc=$?"; tail -2 gpurun_out/bench_dfn.err | cut -c1-300
python - <<'PY'
import json
for f in ("bench_pspnet", "bench_psanet", "bench_dfn"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "eager", round(d["config"]["eager_ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), d["config"]["step_launch"])
    except Exception as e: print(f, "ERR", e)
PY
