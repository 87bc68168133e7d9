#!/usr/bin/env python
"""Whole-sample() time of the c2 / c3 workloads WITHOUT per-launch CUDA events (bench.py brackets every launch with
events, which keeps consecutive kernels from overlapping programmatically): run with DPM_PDL=1 and DPM_PDL=0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import refcheck as R  # noqa: E402
from bench import DT, WORKLOADS  # noqa: E402
import dpm_solver_b200 as new  # noqa: E402

for name in ("c2", "c3"):
    w = WORKLOADS[name]
    x, banks = R.synthetic(w, w["shape"][0], "cuda:0", DT[w["dtype"]])
    s, _ = R._solver(new, w, banks, "cuda:0", state_dtype=DT[w["dtype"]])
    kw = R.sample_kwargs(w)
    for _ in range(5):
        s.sample(x, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s.sample(x, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"DPM_PDL={os.environ.get('DPM_PDL', '1')} {name}: {ms:.4f} ms per sample()  {x.numel() * w['steps'] / ms / 1e6:.1f} GElem/s", flush=True)
