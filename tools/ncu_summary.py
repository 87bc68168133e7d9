#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into a small JSON + markdown table.

    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep --out profiles/r01_ncu_summary
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__cluster_size", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (vals[i], units[i])
        res.append(d)
    return res


def sass_mnemonics(rep, pats=("UBLKCP", "UTMALDG", "UTMASTG", "LDG.E", "STG.E", "LDS", "STS", "SYNCS", "ATOMS", "UCGABAR", "MUFU", "FCHK")):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    cnt = {}
    for line in out.splitlines():
        for p in pats:
            if p in line:
                cnt[p] = cnt.get(p, 0) + 1
    return cnt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    allr = {}
    md = ["| capture | kernel | time µs | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of ncu peak | issue-active % | warp-inst | regs | grid×block | dyn smem |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for rep in a.reps:
        name = os.path.basename(rep).replace(".ncu-rep", "")
        rs = raw(rep)
        allr[name] = {"kernels": rs, "sass_lines": sass_mnemonics(rep)}
        for d in rs:
            g = lambda k: float(d[k][0].replace(",", "")) if k in d else float("nan")
            t = g("gpu__time_duration.sum")
            unit = d.get("gpu__time_duration.sum", ("", "us"))[1]
            t_us = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(unit.replace("second", "s").replace("usecond", "us"), t)
            if unit.startswith("ns"):
                t_us = t / 1e3
            elif unit.startswith("ms"):
                t_us = t * 1e3
            rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
            scale = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(d.get(k, ("", "byte"))[1], 1.0)
            rdb, wrb = rd * scale("dram__bytes_read.sum"), wr * scale("dram__bytes_write.sum")
            kn = d["kernel"].split("(")[0].replace("void dpm::", "")
            md.append(f"| {name} | `{kn}` | {t_us:.1f} | {rdb / 1e6:.1f} | {wrb / 1e6:.1f} | {(rdb + wrb) / t_us / 1e3:.0f} | "
                      f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                      f"{g('smsp__inst_executed.sum'):.3g} | {g('launch__registers_per_thread'):.0f} | {g('launch__grid_size'):.0f}×{g('launch__block_size'):.0f} | "
                      f"{d.get('launch__shared_mem_per_block_dynamic', ('0', ''))[0]} {d.get('launch__shared_mem_per_block_dynamic', ('', ''))[1]} |")
    json.dump(allr, open(a.out + ".json", "w"), indent=1)
    open(a.out + ".md", "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
