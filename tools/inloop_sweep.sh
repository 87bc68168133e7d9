#!/bin/bash
# In-loop launch-shape sweep of the TMA ring (run under gpurun): the isolated kernel probe and the sampling
# loop disagree on the best shape for 5-6 stream kernels (profiles/r01_tma_units.txt vs bench.py), so the
# shapes are compared where they are used -- inside bench.py's timed loop. One line per configuration:
#   workload threads ctas  GElem/s  ms/step  {kernel: GB/s ...}
# Usage: bash tools/inloop_sweep.sh [workloads...]   (default: c2 c3)
mkdir -p gpurun_out
OUT=gpurun_out/inloop_sweep.txt
: > $OUT
WL=${@:-c2 c3}
for w in $WL; do
  for cfg in "0 0" "256 1" "256 2" "256 3" "128 3" "128 4" "128 6" "512 1" "64 8"; do
    set -- $cfg
    line=$(timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-extras --threads $1 --ctas $2 2>/dev/null | tail -1)
    python - "$w" "$1" "$2" "$line" >> $OUT <<'PY'
import json, sys
w, t, c, line = sys.argv[1:5]
try:
    d = json.loads(line)
    ks = {(k.split("|")[0] + "/" + k.split("|")[1][-1] + k.split("|")[2][-1]) if k.count("|") == 2 else k: round(v["gbs"]) for k, v in d["kernels"].items()}
    print(f"{w} threads={t} ctas={c}  {d['value']:.1f} GElem/s  {d['ms_per_step']:.3f} ms  {ks}")
except Exception as e:
    print(f"{w} threads={t} ctas={c}  FAILED {e}: {line[:120]}")
PY
  done
done
cat $OUT
