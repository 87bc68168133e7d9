"""T5 drop-in acceptance: the reference's own Stable-Diffusion adapter (DPMSolverSampler,
examples/stable-diffusion/ldm/models/diffusion/dpm_solver/sampler.py) is executed UNMODIFIED twice --
once on top of its vendored copy of the solver, once on top of dpm_solver_b200 -- with a stub
LatentDiffusion model. Outputs (sample, every intermediate, stochastic_encode) must be bit-identical.
Needs /root/reference (build container only); skipped elsewhere."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("DPM_REFERENCE", "/root/reference")
ADAPTER_DIR = os.path.join(REF, "examples", "stable-diffusion", "ldm", "models", "diffusion", "dpm_solver")
pytestmark = pytest.mark.skipif(not os.path.isdir(ADAPTER_DIR), reason="reference tree not available")


def load_adapter(pkg_name, solver_module):
    """Import sampler.py inside a synthetic package whose `.dpm_solver` is `solver_module`."""
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = []
    sys.modules[pkg_name] = pkg
    sys.modules[pkg_name + ".dpm_solver"] = solver_module
    spec = importlib.util.spec_from_file_location(pkg_name + ".sampler", os.path.join(ADAPTER_DIR, "sampler.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[pkg_name + ".sampler"] = mod
    spec.loader.exec_module(mod)
    # the adapter pins its buffers to "cuda"; the acceptance run is on CPU
    mod.DPMSolverSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    return mod


class StubLatentDiffusion:
    """What DPMSolverSampler touches: alphas_cumprod, betas.device, device, apply_model."""

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
        self.betas = betas.float()
        self.alphas_cumprod = torch.cumprod(1 - betas, 0).float()
        self.device = torch.device("cpu")
        self.calls = []

    def apply_model(self, x, t, c):
        self.calls.append((float(t[0]), tuple(x.shape)))
        return 0.1 * x + ((t * 0.001) * 0.05 - 0.02).reshape(-1, 1, 1, 1) + 0.05 * c.reshape(-1, 1, 1, 1)


def vendored_solver():
    spec = importlib.util.spec_from_file_location("_ref_sd_dpm_solver", os.path.join(ADAPTER_DIR, "dpm_solver.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("order,steps,method", [(2, 20, "multistep"), (3, 12, "multistep"), (2, 9, "singlestep")])
def test_stable_diffusion_adapter_runs_unchanged(oracle_backend, order, steps, method):
    import dpm_solver_b200
    ref_mod = load_adapter("_adapter_ref", vendored_solver())
    new_mod = load_adapter("_adapter_b200", dpm_solver_b200)
    B, shape = 2, (4, 16, 16)
    x_T = torch.randn(B, *shape, generator=torch.Generator().manual_seed(3))
    cond, uncond = torch.ones(B, 1), torch.zeros(B, 1)
    outs = []
    for mod in (ref_mod, new_mod):
        model = StubLatentDiffusion()
        sampler = mod.DPMSolverSampler(model)
        assert sampler.noise_schedule.total_N == 1000
        x, inter = sampler.sample(S=steps, batch_size=B, shape=shape, conditioning=cond, x_T=x_T.clone(),
                                  unconditional_guidance_scale=7.5, unconditional_conditioning=uncond,
                                  order=order, method=method, verbose=False)
        enc = sampler.stochastic_encode(x_T, 0.5, noise=torch.ones_like(x_T))
        outs.append((x, inter, enc, model.calls))
    (xr, ir, er, cr), (xn, in_, en, cn) = outs
    assert cr == cn                                    # same network calls: doubled batch, same time labels
    np.testing.assert_array_equal(xn.numpy(), xr.numpy())
    assert len(ir) == len(in_)
    for a, b in zip(ir, in_):
        np.testing.assert_array_equal(b.numpy(), a.numpy())
    np.testing.assert_array_equal(en.numpy(), er.numpy())


def test_root_module_name_is_a_drop_in():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.modules.pop("dpm_solver_pytorch", None)
    import dpm_solver_pytorch as m
    import dpm_solver_b200
    assert m.DPM_Solver is dpm_solver_b200.DPM_Solver and m.NoiseScheduleVP is dpm_solver_b200.NoiseScheduleVP
    assert m.model_wrapper is dpm_solver_b200.model_wrapper
