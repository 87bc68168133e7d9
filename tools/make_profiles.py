#!/usr/bin/env python
"""Collect a GPU round's evidence from gpurun_out/ into profiles/ (tracked): bench JSON lines, ncu
launch lists and their per-kernel shares next to the CUDA-event shares of the same workload, the
`ncu --set full` summaries, the variant sweep, and roofline_traffic.json (read by bench.py).

    python tools/make_profiles.py --round r01
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[iid], {"k": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
    return list(d.values())


def short(k):
    return k.replace("void ", "").replace("dpm::", "").split("(")[0]


def agg_launches(ls):
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for v in ls:
        a = agg[short(v["k"])]
        a[0] += 1
        a[1] += v["gpu__time_duration.sum"]
        a[2] += v.get("dram__bytes_read.sum", 0.0)
        a[3] += v.get("dram__bytes_write.sum", 0.0)
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    a = ap.parse_args()
    R = a.round
    os.makedirs(DST, exist_ok=True)
    md = [f"# profiles — round {R[1:]}", "",
          "All numbers measured on one B200 (sm_100a) of this pool through `gpurun`; HBM peak denominator = "
          "`MEASURED_PEAKS.json` `hbm_gbs` = 6574.5 GB/s (torch copy, read+write). Event timings come from "
          "`bench.py` (CUDA events on the launching stream, inputs > L2, >= 3 warm-ups); ncu timings are "
          "cold-cache and serialised — compare shares, not absolutes. Regenerate with "
          "`bash tools/gpu_run.sh tests_all smoke bench bench_each ref launches ncu_step ncu_quantile latency` under gpurun, "
          f"then `python tools/make_profiles.py --round {R}`. Earlier rounds: `profiles/r01_README.md`.", ""]
    traffic = {}
    # ---- bench lines ----
    md += ["## bench.py lines", "", "| workload | GElem/s | ms/step | HBM GB/s (all launches) | dominant kernel | achieved GB/s | frac of peak | e2e GElem/s | launches | SM MHz |", "|---|---|---|---|---|---|---|---|---|---|"]
    benches = {}
    for w in ("c2", "c3", "c4"):
        p = os.path.join(SRC, f"bench_{w}.json")
        if w == "c2" and os.path.exists(os.path.join(SRC, "bench_default.json")):
            p = os.path.join(SRC, "bench_default.json")       # the driver's command: carries workloads / kernels_alone / baselines
        if not os.path.exists(p):
            continue
        lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if not lines:
            continue
        d = json.loads(lines[-1])
        benches[w] = d
        json.dump(d, open(os.path.join(DST, f"{R}_bench_{w}.json"), "w"), indent=1)
        r = d["roofline"]
        md.append(f"| {w}: {d['config']['workload']} | {d['value']:.1f} | {d['ms_per_step']:.3f} | {d['hbm_gbs_total']:.0f} | `{r['kernel']}` | "
                  f"{r['achieved']:.0f} | {r['frac']:.3f} | {d['e2e']['value']:.1f} | {d['gpu_launches']} | {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')} |")
    p = os.path.join(SRC, "bench_ref.json")
    if os.path.exists(p):
        lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if lines:
            d = json.loads(lines[-1])
            json.dump(d, open(os.path.join(DST, f"{R}_bench_reference_arm.json"), "w"), indent=1)
            cb = d["cpu_baseline"]
            what = "the UNMODIFIED reference from oracle/_ref on CPU tensors" if cb.get("kind") == "reference" else "oracle port in its torch-CPU namespace"
            md += ["", f"Reference arm (`bench.py --impl reference`, {what}): **{d['value']:.3f} GElem/s** on "
                   f"{cb['cores']} host threads (host has {cb.get('host_cores')} cores, {cb.get('usable_cores')} usable by the job), sample {cb['sample']}."]
    if "c2" in benches and "kernels_alone" in benches["c2"]:
        md += ["", "North-star kernel timed alone (fused 3rd-order multistep update, x + 3 buffers -> x_t, `[4096,4,64,64]`, rotating through 3 buffer sets):", "",
               "| kernel | bytes/launch | median µs | GB/s | frac of measured peak | GElem/s |", "|---|---|---|---|---|---|"]
        for k, v in benches["c2"]["kernels_alone"].items():
            md.append(f"| {k} | {v['bytes_per_launch']} | {v['median_us']:.1f} | {v['gbs']:.0f} | {v['frac_of_peak']:.3f} | {v['gelem_s']:.0f} |")
        cb = benches["c2"].get("cpu_baseline")
        if cb:
            md += ["", f"CPU baseline beside it (same job, rank 0, kind `{cb.get('kind')}`): {cb['value']:.3f} GElem/s, {cb['cores']} threads "
                   f"of {cb.get('host_cores')} host cores, {cb['sample']}."]
        eg = benches["c2"].get("eager_cuda_baseline")
        if eg and "value" in eg:
            md += ["", f"Second baseline, same GPU: the reference (kind `{eg.get('kind')}`: `reference` = the unmodified file from oracle/_ref) as stock "
                   f"eager PyTorch CUDA kernels, fp32 state, full C2 shape: **{eg['value']:.1f} GElem/s** ({eg['ms_per_step']:.1f} ms per 20-step sample()) — "
                   f"{benches['c2']['value'] / eg['value']:.0f}x below the fused bf16 path ({benches['c2']['value']:.0f} GElem/s)."]
    for w, d in benches.items():
        par = d.get("parity") or {}
        md += ["", f"{w} parity (outside the timed region, {par.get('rows')} rows of the run's own output vs {par.get('reference')}): "
               f"checked={par.get('parity_checked')}, max rel err {par.get('max_rel_err')}, rms rel err {par.get('rms_rel_err')}, "
               f"bit_exact={par.get('bit_exact')}, bound={par.get('bound')}."]
    if "c2" in benches and "workloads" in benches["c2"]:
        md += ["", "The default `python bench.py` line (what the driver records) also carries the other single-GPU configs under `workloads`:", "",
               "| workload | GElem/s | ms/step | dominant kernel | GB/s | frac | parity_checked |", "|---|---|---|---|---|---|---|"]
        for n, o in benches["c2"]["workloads"].items():
            if "error" in o:
                md.append(f"| {n} | error: {o['error']} | | | | | |")
            else:
                md.append(f"| {n} | {o['value']:.1f} | {o['ms_per_step']:.3f} | `{o['roofline']['kernel']}` | {o['roofline']['achieved']:.0f} | {o['roofline']['frac']:.3f} | {o['config'].get('parity_checked')} |")
    md.append("")
    # ---- per-kernel event table + ncu launch share ----
    for w, d in benches.items():
        md += [f"## {w}: per-kernel view", "", "CUDA events inside the timed region (bench.py):", "",
               "| kernel key | launches | avg µs | algorithmic bytes/launch | GB/s | share of kernel time |", "|---|---|---|---|---|---|"]
        tot = sum(v["ms"] for v in d["kernels"].values())
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
            md.append(f"| `{k}` | {v['launches']} | {v['avg_us']:.1f} | {v['bytes_per_launch']:.0f} | {v['gbs']:.0f} | {v['ms'] / tot:.1%} |")
        lp = os.path.join(SRC, f"launches_{w}.csv")
        if os.path.exists(lp):
            shutil.copy(lp, os.path.join(DST, f"{R}_launches_{w}.csv"))
            ls = launch_list(lp)
            agg = agg_launches(ls)
            tot = sum(x[1] for x in agg.values())
            md += ["", f"ncu launch list of the same command (`profiles/{R}_launches_{w}.csv`, `--metrics gpu__time_duration.sum,dram__bytes_*`, `--clock-control none`):", "",
                   "| kernel | launches | avg µs | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
            traffic[w] = {}
            for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                md.append(f"| `{k}` | {x[0]} | {x[1] / x[0] / 1e3:.1f} | {x[1] / tot:.1%} | {x[2] / x[0] / 1e6:.1f} | {x[3] / x[0] / 1e6:.1f} |")
            # traffic of the dominant kernel: ncu launches of the matching kernel with the largest time share
            dom_k, dom = max(((k, x) for k, x in agg.items() if k.startswith("k_step") or k.startswith("k_q")), key=lambda kv: kv[1][1])
            traffic[w][d["roofline"]["kernel"]] = (dom[2] + dom[3]) / dom[0]
            traffic[w]["_ncu_kernel"] = dom_k
        md.append("")
    if traffic:
        traffic["_source"] = f"profiles/{R}_launches_<workload>.csv"
        json.dump(traffic, open(os.path.join(DST, "roofline_traffic.json"), "w"), indent=1)
    # ---- ncu full ----
    reps = sorted(f for f in os.listdir(SRC) if f.endswith(".ncu-rep"))
    pre = os.path.join(SRC, "ncu_summary.md")
    if os.path.exists(pre) and os.path.exists(os.path.join(SRC, "ncu_summary.json")):
        # tools/gpu_run.sh summarised the captures on the GPU box (the .ncu-rep files stay there)
        # merge by capture name: a partial round (e.g. only the quantile kernels re-captured) keeps the rest
        dmd, djs = os.path.join(DST, f"{R}_ncu_summary.md"), os.path.join(DST, f"{R}_ncu_summary.json")
        rows = OrderedDict()
        head = []
        for path in (dmd, pre):
            if os.path.exists(path):
                lines = open(path).read().strip().splitlines()
                head = lines[:2]
                for l in lines[2:]:
                    rows[l.split("|")[1].strip()] = l
        open(dmd, "w").write("\n".join(head + [rows[k] for k in sorted(rows)]) + "\n")
        js = json.load(open(djs)) if os.path.exists(djs) else {}
        js.update(json.load(open(os.path.join(SRC, "ncu_summary.json"))))
        json.dump(js, open(djs, "w"), indent=1)
        md += ["## `ncu --set full` captures (one launch each, `--clock-control none`)", "", open(dmd).read().strip(), "",
               f"Full metric dump: `profiles/{R}_ncu_summary.json` (incl. stall reasons and SASS mnemonic counts: `UBLKCP` = TMA bulk copy, "
               "`SYNCS` = mbarrier, `LDG.E…256/128` vector loads, `UCGABAR` = cluster barrier).", ""]
    elif reps:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] + [os.path.join(SRC, f) for f in reps] +
                             ["--out", os.path.join(DST, f"{R}_ncu_summary")], capture_output=True, text=True).stdout
        md += ["## `ncu --set full` captures (one launch each, `--clock-control none`)", "", out.strip(), "",
               f"Full metric dump: `profiles/{R}_ncu_summary.json` (incl. stall reasons and SASS mnemonic counts: `UBLKCP` = TMA bulk copy, "
               "`SYNCS` = mbarrier, `LDG.E…256/128` vector loads, `UCGABAR` = cluster barrier).", ""]
    # ---- sweeps ----
    for name, title in (("sweep2.jsonl", "variant / threads / CTAs-per-SM sweep (kernel alone, rotating buffers)"),
                        ("chain.jsonl", "same kernels with the sampling loop's data flow (x_{i+1} = out_i, m1_{i+1} = m_out_i)")):
        p = os.path.join(SRC, name)
        if not os.path.exists(p):
            continue
        rows = [json.loads(l) for l in open(p) if l.startswith("{")]
        shutil.copy(p, os.path.join(DST, f"{R}_{name}"))
        g = defaultdict(list)
        for r in rows:
            g[(r["form"], r.get("n_model", 1), r["dtype"])].append(r)
        md += [f"## {title}", "", "| form | n_model | dtype | best direct (threads×CTAs/SM) GB/s | best TMA ring GB/s | TMA 256×2 GB/s |", "|---|---|---|---|---|---|"]
        for k, rs in g.items():
            bd = max((r for r in rs if r["variant"] == 0), key=lambda r: r["gbs"], default=None)
            bt = max((r for r in rs if r["variant"] == 1), key=lambda r: r["gbs"], default=None)
            t22 = [r for r in rs if r["variant"] == 1 and r["threads"] == 256 and r["ctas"] == 2]
            f = lambda r: f"{r['gbs']:.0f} ({r['threads']}×{r['ctas']})" if r else "-"
            md.append(f"| {k[0]} | {k[1]} | {k[2]} | {f(bd)} | {f(bt)} | {t22[0]['gbs']:.0f} |" if t22 else f"| {k[0]} | {k[1]} | {k[2]} | {f(bd)} | {f(bt)} | - |")
        md.append("")
    for extra in ("host_overhead.txt", "inloop_sweep.txt", "l2_group_probe.txt"):
        p = os.path.join(SRC, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, f"{R}_{extra}"))
    appendix = os.path.join(DST, f"{R}_appendix.md")
    if os.path.exists(appendix):
        md.append(open(appendix).read().rstrip())
    open(os.path.join(DST, "README.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md)[:3000])


if __name__ == "__main__":
    main()
