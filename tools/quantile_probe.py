#!/usr/bin/env python
"""Which path does the quantile pipeline take, and how long does each kernel need?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dpm_solver_b200 import ops
from dpm_solver_b200.ops import StepArgs

be = ops.CudaBackend()
B, ps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 3 * 256 * 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B * ps, device="cuda", generator=g)
e = torch.randn(B * ps, device="cuda", generator=g)
for alpha, sigma in ((0.0063, 0.99998), (0.5, 0.866), (0.999, 0.04)):
    a = StepArgs(form=0, n_model=1, e_cond=e, xe=x, predict_x0=True, alpha_e=alpha, sigma_e=sigma,
                 per_sample=ps, state_dtype=torch.float32)
    s, hdr = be.dynamic_threshold(a, 0.995, 1.0, return_stats=True)
    torch.cuda.synchronize()
    h = hdr.cpu()
    path = h[:, 4]
    lo = int(0.995 * (ps - 1))
    print(f"alpha={alpha}: bracket={int((path == 1).sum())} fallback={int((path == 2).sum())} "
          f"C_lt[min,max]=({int(h[:, 2].min())},{int(h[:, 2].max())}) C_in[min,max]=({int(h[:, 3].min())},{int(h[:, 3].max())}) target rank {lo}")
    bad = (path == 2).nonzero().flatten()[:5]
    for b in bad.tolist():
        print("   fallback sample", b, "lo_key", hex(int(h[b, 0]) & 0xffffffff), "hi_key", hex(int(h[b, 1]) & 0xffffffff), "C_lt", int(h[b, 2]), "C_in", int(h[b, 3]))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        be.dynamic_threshold(a, 0.995, 1.0)
    t1.record()
    torch.cuda.synchronize()
    print(f"   {t0.elapsed_time(t1) / 5 * 1e3:.1f} us per call")
