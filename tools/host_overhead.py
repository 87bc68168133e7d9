#!/usr/bin/env python
"""Small-batch (latency) regime: the reference's real call sites run B <= 16 (scripts/txt2img.py:303), where the
kernels take a few microseconds and the HOST path per solver step is what a user waits for.

For B in {1, 8, 64} at [B,4,64,64], DPM-Solver++ 2M, 20 steps, a network that returns a stored tensor:
  * product, steady state (prepared launches, ops.PreparedStep)   us per solver step, wall clock
  * product with the prepared path disabled (general path)        us per solver step
  * product, whole loop captured in one CUDA graph (capture())    us per solver step
  * the UNMODIFIED reference (oracle/_ref) as eager CUDA ops on the same GPU
and a cProfile of the steady-state loop at B = 8.

    python tools/host_overhead.py
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from cases import make_betas  # noqa: E402
from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper  # noqa: E402

STEPS = 20
betas = torch.from_numpy(make_betas("sd")[1])


def wall(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n / STEPS * 1e6


ref = None
try:
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load("dpm_solver_pytorch")
except Exception:
    pass

print(f"{'B':>4} {'prepared':>10} {'general':>10} {'graph':>10} {'reference eager CUDA':>22}   (us per solver step, DPM-Solver++ 2M x {STEPS})")
prof_target = None
for B in (1, 8, 64):
    for dt in (torch.float32, torch.bfloat16):
        ns = NoiseScheduleVP("discrete", betas=betas)
        x = torch.randn(B, 4, 64, 64, device="cuda").to(dt)
        bank = torch.randn(B, 4, 64, 64, device="cuda").to(dt)
        s = DPM_Solver(model_wrapper(lambda xx, tt: bank, ns), ns, state_dtype=None if dt == torch.float32 else dt)
        run = lambda: s.sample(x, steps=STEPS, order=2)
        t_prep = wall(run, 200)

        def general():
            s._prep_cache.clear()
            return s.sample(x, steps=STEPS, order=2)
        t_gen = wall(general, 100)
        g = s.capture(x, steps=STEPS, order=2)
        t_graph = wall(lambda: g(x), 200)
        t_ref = float("nan")
        if ref is not None and dt == torch.float32:
            nr = ref.NoiseScheduleVP("discrete", betas=betas)
            sr = ref.DPM_Solver(ref.model_wrapper(lambda xx, tt: bank, nr), nr, algorithm_type="dpmsolver++")
            t_ref = wall(lambda: sr.sample(x, steps=STEPS, order=2), 20)
        print(f"{B:>4} {t_prep:>10.1f} {t_gen:>10.1f} {t_graph:>10.1f} {t_ref:>22.1f}   {str(dt).replace('torch.', '')}", flush=True)
        if B == 8 and dt == torch.float32:
            prof_target = run

pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    prof_target()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
