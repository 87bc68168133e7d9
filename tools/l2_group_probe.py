#!/usr/bin/env python
"""How much of the thresholded step's second read of (x, eps) can the 126 MB L2 serve?

For a group of G samples of [3,256,256] fp32 (1.57 MB of (x, eps) per sample): flush L2, run the quantile on
the group, then the fused 3M step on the SAME group, and time the step with CUDA events. Compared with the
step on a cold L2, the saving is what a group-ordered (L2-aware) schedule can win. Under ncu
(--cache-control none --metrics dram__bytes_read.sum) the per-kernel DRAM bytes show the hit rate directly.

    python tools/l2_group_probe.py [G ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dpm_solver_b200 import ops  # noqa: E402
from dpm_solver_b200.ops import StepArgs  # noqa: E402

be = ops.CudaBackend()
ps = 3 * 256 * 256
Gs = [int(v) for v in sys.argv[1:]] or [8, 16, 24, 32, 48, 64, 96]
flush = torch.zeros(256 << 20, dtype=torch.float32, device="cuda")     # 1 GiB, evicted by READING it: no dirty lines left behind
sink = torch.zeros(1, device="cuda")


def evict():
    sink.copy_(flush.sum().reshape(1))
for G in Gs:
    n = G * ps
    mk = lambda: torch.randn(n, device="cuda")
    x, e, m1, m2 = mk(), mk(), mk(), mk()
    out, mo = torch.empty_like(x), torch.empty_like(x)

    def args(thr=None):
        return StepArgs(form=5, n_model=1, x=x, xe=x, e_cond=e, m1=m1, m2=m2, predict_x0=True, alpha_e=0.83, sigma_e=0.55,
                        a=0.95, c0=-0.1, c1=0.05, c2=-0.01, w0=1.02, w1=0.98, w2=0.51, w3=0.5, want_m_out=True,
                        state_dtype=torch.float32, per_sample=ps, thr=thr, out=out, m_out=mo)

    res = {}
    for mode in ("cold", "after_quantile"):
        ts = []
        for rep in range(5):
            evict()
            thr = be.dynamic_threshold(args(), 0.995, 1.0)
            if mode == "cold":
                evict()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.step(args(thr))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[mode] = sorted(ts)[len(ts) // 2]
    byts = 6 * n * 4
    print(f"G={G:4d} (x,eps)={2 * n * 4 / 1e6:6.1f} MB  step cold {res['cold']:7.1f} us ({byts / res['cold'] / 1e3:6.0f} GB/s)  "
          f"after quantile {res['after_quantile']:7.1f} us ({byts / res['after_quantile'] / 1e3:6.0f} GB/s alg.)  "
          f"saving {100 * (1 - res['after_quantile'] / res['cold']):.0f}%", flush=True)
