#!/usr/bin/env python
"""How far apart do adaptive solves land when only the LAST ULP of the schedule scalars differs?

For random adaptive configurations (tests/test_adaptive.py::_adaptive_cfgs) run on the GPU box:
  ref-cpu   : the unmodified reference on CPU tensors (oracle/_ref)                    -- the yardstick
  ref-cuda  : the unmodified reference on CUDA tensors (its scalars from the device's libm)
  host-ctl  : dpm_solver_b200, controller on the host (scalars identical to ref-cpu)
  dev-ctl   : dpm_solver_b200, controller on the device (csrc/adaptive_ctl.cu)
and print NFE and max|y - y_ref-cpu| / max|y_ref-cpu| for each.
"""
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import dpm_solver_b200 as new  # noqa: E402
from helpers import rel_err  # noqa: E402
from oracle import ref_loader  # noqa: E402
from test_adaptive import _adaptive_cfgs  # noqa: E402
from test_random_configs_vs_reference import run_wide  # noqa: E402

ref = ref_loader.load("dpm_solver_pytorch")


def on_gpu(mod):
    class M:
        NoiseScheduleVP = mod.NoiseScheduleVP

        @staticmethod
        def model_wrapper(net, ns, **kw):
            return mod.model_wrapper(net, ns, **{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()})

        class DPM_Solver(mod.DPM_Solver):
            def sample(self, x, **kw):
                return super().sample(x.cuda(), **kw).cpu()
    return M


def run(mod, c):
    with mock.patch("builtins.print") as pr:
        y, _, _ = run_wide(mod, c)
    return y, pr.call_args[0][-1]


print(f"{'cfg':>4} {'order':>5} {'type':>9} {'algo':>11} {'model':>7} {'cfg':>4} {'nfe ref/cuda/host/dev':>22} {'ref-cuda':>10} {'host-ctl':>10} {'dev-ctl':>10}")
k = 0
for chunk in range(4):
    for c in _adaptive_cfgs(10, 9000 + chunk):
        yr, nr = run(ref, c)
        if not torch.isfinite(yr).all():
            continue
        yc, nc = run(on_gpu(ref), c)
        new.DPM_Solver.adaptive_controller = "host"
        yh, nh = run(on_gpu(new), c)
        new.DPM_Solver.adaptive_controller = "device"
        yd, nd = run(on_gpu(new), c)
        e = [rel_err(v.numpy(), yr.numpy()) for v in (yc, yh, yd)]
        print(f"{k:>4} {c['order']:>5} {c['solver_type']:>9} {c['algo']:>11} {c['model_type']:>7} {str(c['cfg']):>4} "
              f"{f'{nr}/{nc}/{nh}/{nd}':>22} {e[0]:>10.2e} {e[1]:>10.2e} {e[2]:>10.2e}", flush=True)
        k += 1
