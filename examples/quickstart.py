#!/usr/bin/env python
"""Minimal end-to-end use on a CUDA device: the reference's three-call recipe (README.md:376-568 of the
reference), unchanged, on top of dpm_solver_b200 -- with a toy conv network standing in for the U-Net.

    python examples/quickstart.py [--batch 64] [--steps 20] [--bf16] [--graph]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dpm_solver_pytorch import DPM_Solver, NoiseScheduleVP, model_wrapper  # noqa: E402  (drop-in module name)


class ToyEps(torch.nn.Module):
    """A stand-in for a text-conditioned U-Net: eps(x, t, cond)."""

    def __init__(self, ch=4):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Conv2d(ch, 32, 3, padding=1), torch.nn.SiLU(), torch.nn.Conv2d(32, ch, 3, padding=1))

    def forward(self, x, t, cond):
        return self.net(x.float()).to(x.dtype) + 0.01 * torch.sin(t * 1e-3).reshape(-1, 1, 1, 1).to(x.dtype) * cond.reshape(-1, 1, 1, 1).to(x.dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--bf16", action="store_true", help="keep the solver state in bf16 storage")
    ap.add_argument("--graph", action="store_true", help="capture the whole sampling loop in one CUDA graph")
    a = ap.parse_args()
    dev = "cuda"
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2      # SD-v1 schedule
    ns = NoiseScheduleVP("discrete", betas=betas)
    unet = ToyEps().to(dev).eval()
    cond, uncond = torch.ones(a.batch, 1, device=dev), torch.zeros(a.batch, 1, device=dev)
    model_fn = model_wrapper(unet, ns, model_type="noise", guidance_type="classifier-free", condition=cond,
                             unconditional_condition=uncond, guidance_scale=7.5)
    solver = DPM_Solver(model_fn, ns, algorithm_type="dpmsolver++", state_dtype=torch.bfloat16 if a.bf16 else None)
    x_T = torch.randn(a.batch, 4, 64, 64, device=dev)
    kw = dict(steps=a.steps, order=2, skip_type="time_uniform", method="multistep")
    with torch.no_grad():
        run = solver.capture(x_T, **kw) if a.graph else (lambda x: solver.sample(x, **kw))
        run(x_T)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            x0 = run(x_T)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"sampled {tuple(x0.shape)} {x0.dtype} in {dt * 1e3:.2f} ms per call ({a.steps} NFE, CFG 7.5, graph={a.graph})  |x0| mean {float(x0.float().abs().mean()):.4f}")


if __name__ == "__main__":
    main()
