#!/bin/bash
mkdir -p gpurun_out
for v in d e f; do VAR=$v timeout 300 python tools/diag_graph2.py 2>&1 | grep -v "Warning\|detach\|l = gs" | tail -2; done
timeout 900 python -m pytest tests/test_gpu_bisenet.py -q -m gpu -p no:cacheprovider -k "graphed or mirror" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
for f in ("bench_a",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", "eager", round(d["config"]["eager_ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "launches", d["gpu_launches"], d["config"]["step_launch"], d.get("cpu_baseline"), d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
