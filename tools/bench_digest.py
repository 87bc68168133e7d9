#!/usr/bin/env python
"""Print the few numbers of a bench.py JSON line a human wants to see first."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(path, "FAILED", e)
        continue
    if d.get("impl") == "reference":
        print(path, "reference arm", round(d["value"], 4), d["unit"], d["cpu_baseline"])
        continue

    def line(tag, o):
        r = o["roofline"]
        par = o.get("parity") or {}
        print(f"{tag}: {o['value']:.1f} GElem/s  {o['ms_per_step']:.3f} ms/step (instrumented {o.get('ms_per_step_instrumented', 0):.3f})  dom={r['kernel']} {r['achieved']:.0f} GB/s frac={r['frac']:.3f} "
              f"total={o['hbm_gbs_total']:.0f} GB/s  parity={par.get('parity_checked')} max_rel={par.get('max_rel_err')}  clocks={o['clocks']}")
        for k, v in sorted(o["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
            print(f"    {k:44s} n={v['launches']:4d} avg={v['avg_us']:8.1f} us  {v['gbs']:.0f} GB/s")
    line(path, d)
    if "e2e" in d:
        print("  e2e:", {k: v for k, v in d["e2e"].items() if k != "limiter"})
    for n, o in (d.get("workloads") or {}).items():
        if "error" in o:
            print("  ", n, o)
        else:
            line("  " + n, o)
    for k in ("kernels_alone", "cpu_baseline", "eager_cuda_baseline"):
        if k in d:
            print("  ", k, d[k])
    if "shard_parity" in d.get("config", {}):
        print("  shard_parity:", d["config"]["shard_parity"])
