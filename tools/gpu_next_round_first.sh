#!/bin/bash
# First GPU call of the next round (run under gpurun, ~12 min): what this round staged without hardware.
#  1. full GPU suite with the staged reference-rounding tests reported explicitly (XPASS = ready to unmark)
#  2. the in-loop launch-shape sweep of the TMA ring for c2 / c3
#  3. bench lines of the three workloads for a same-box baseline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; tail -50 gpurun_out/pytest_gpu.log | grep -E "XPASS|XFAIL|passed|failed" | tail -50
timeout 900 bash tools/inloop_sweep.sh c2 c3 > /dev/null 2>&1; cat gpurun_out/inloop_sweep.txt
for w in c2 c3 c4; do timeout 600 python bench.py --workload $w --steps 5 --no-extras > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python - $w <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/bench_{w}.json").read().splitlines() if l.startswith("{")][-1])
    print(w, round(d["value"], 1), "GElem/s", round(d["ms_per_step"], 3), "ms", "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(w, "FAILED", e)
PY
done
