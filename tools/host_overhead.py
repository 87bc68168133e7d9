#!/usr/bin/env python
"""Host-side cost of one solver step: run sample() on a tiny batch (kernels take ~2 us, so the
wall clock is the Python + launch path) and print a cProfile of the loop."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from cases import make_betas  # noqa: E402
from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper  # noqa: E402

ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("sd")[1]))
x = torch.randn(8, 4, 64, 64, device="cuda").bfloat16()
bank = torch.randn(8, 4, 64, 64, device="cuda").bfloat16()
s = DPM_Solver(model_wrapper(lambda xx, tt: bank, ns), ns, state_dtype=torch.bfloat16)
for _ in range(5):
    s.sample(x, steps=20, order=2)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    s.sample(x, steps=20, order=2)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / N
print(f"sample(): {dt * 1e3:.3f} ms  -> {dt / 20 * 1e6:.1f} us per solver step (host path)")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    s.sample(x, steps=20, order=2)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
