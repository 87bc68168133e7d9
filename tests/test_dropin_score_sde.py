"""T5, second acceptance target: the reference's score_sde glue (`get_dpm_solver_sampler`,
examples/score_sde_pytorch/sampling.py:505-558) executed UNMODIFIED on top of (a) its vendored solver
copy and (b) dpm_solver_b200 -- continuous 'linear' VP schedule, singlestep order 3, logSNR grid,
optional denoise / thresholding. The model zoo it imports (models.utils, sde_lib) is stubbed.
Needs /root/reference; skipped elsewhere."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("DPM_REFERENCE", "/root/reference")
SDE_DIR = os.path.join(REF, "examples", "score_sde_pytorch")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(SDE_DIR, "sampling.py")), reason="reference tree not available")


def net(x, t):
    return 0.1 * x + ((t * 0.05) - 0.02).reshape(-1, 1, 1, 1)


def load_sampling(solver_module, tag):
    mutils = types.ModuleType("models.utils")
    mutils.from_flattened_numpy = mutils.to_flattened_numpy = mutils.get_score_fn = lambda *a, **k: None
    mutils.get_noise_fn = lambda sde, model, train=False, continuous=True: (lambda x, t: model(x, t))
    models = types.ModuleType("models")
    models.utils = mutils
    sde_lib = types.ModuleType("sde_lib")
    saved = {k: sys.modules.get(k) for k in ("models", "models.utils", "sde_lib", "dpm_solver")}
    sys.modules.update({"models": models, "models.utils": mutils, "sde_lib": sde_lib, "dpm_solver": solver_module})
    try:
        spec = importlib.util.spec_from_file_location("_score_sde_sampling_" + tag, os.path.join(SDE_DIR, "sampling.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class StubVPSDE:
    beta_0, beta_1, T = 0.1, 20.0, 1.0

    def prior_sampling(self, shape):
        return torch.randn(*shape, generator=torch.Generator().manual_seed(21))


def vendored():
    """The reference solver the glue runs on. The example's own vendored (older) copy calls
    `correcting_x0_fn(x0)` with one argument and raises TypeError with thresholding
    (examples/score_sde_pytorch/dpm_solver.py:449), so the current root file is used instead."""
    spec = importlib.util.spec_from_file_location("_ref_root_dpm_solver", os.path.join(REF, "dpm_solver_pytorch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kw", [dict(), dict(denoise=True, steps=13), dict(algorithm_type="dpmsolver++", thresholding=True, order=2, steps=8),
                                dict(skip_type="time_uniform", method="multistep", order=2, steps=12)])
def test_score_sde_glue_runs_unchanged(oracle_backend, kw):
    import dpm_solver_b200
    outs = []
    for tag, solver in (("ref", vendored()), ("b200", dpm_solver_b200)):
        sampling = load_sampling(solver, tag)
        fn = sampling.get_dpm_solver_sampler(StubVPSDE(), (2, 3, 8, 8), lambda v: v, device="cpu", **kw)
        x, nfe = fn(net)
        outs.append((x, nfe))
    assert outs[0][1] == outs[1][1]
    np.testing.assert_array_equal(outs[1][0].numpy(), outs[0][0].numpy())
