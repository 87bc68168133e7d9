#!/bin/bash
# bench.py inside-the-loop numbers for a list of "threads ctas" TMA launch shapes: tools/shape_probe.sh c3 "0 0" "384 1" ...
mkdir -p gpurun_out
w=$1; shift
python -c "import torch" 2>/dev/null
for cfg in "$@"; do
  set -- $cfg
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --threads $1 --ctas $2 > gpurun_out/shape.json 2> gpurun_out/shape.err
  python - "$w" "$1" "$2" <<'PY'
import json, sys
w, t, c = sys.argv[1:4]
try:
    d = json.loads([l for l in open("gpurun_out/shape.json").read().splitlines() if l.startswith("{")][-1])
    ks = {(k.split("|")[0] + "/" + k.split("|")[1][-1] + k.split("|")[2][-1]) if k.count("|") == 2 else k: round(v["gbs"]) for k, v in d["kernels"].items()}
    print(f"{w} threads={t} ctas={c}  {d['value']:.1f} GElem/s  {d['ms_per_step']:.4f} ms  {ks}", flush=True)
except Exception as e:
    print(f"{w} threads={t} ctas={c} FAILED {e}", open("gpurun_out/shape.err").read()[-300:], flush=True)
PY
done
