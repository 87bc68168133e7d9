#!/usr/bin/env python
"""Digest of the ncu CSV written by `gpu_run.sh l2ncu`: DRAM bytes of every k_step_direct launch of
tools/l2_group_probe.py (10 launches per G: 5 cold, 5 right after the quantile on the same samples)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_id, i_m, i_v = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = {}
for r in rows[1:]:
    per.setdefault(int(r[i_id]), {})[r[i_m]] = float(r[i_v].replace(",", ""))
ids = sorted(per)
for k, i in enumerate(ids):
    m = per[i]
    print(f"launch {k:3d}: time {m.get('gpu__time_duration.sum', 0) / 1e3:7.1f} us  dram read {m.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB  "
          f"write {m.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB  L2 hit {m.get('lts__t_sector_hit_rate.pct', 0):5.1f} %")
