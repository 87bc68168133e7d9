"""dpm_solver_adaptive (:956-1010): same NFE and the same sample as the reference. The error estimate
is a reduction (its summation order differs between torch-CPU, numpy and the CUDA kernel), so the
sample is compared with the north-star tolerance instead of bit for bit; the accept/reject decisions
-- hence the NFE the reference prints -- must be identical."""
import numpy as np
import pytest
import torch

from cases import exact_net, make_betas, seeded
from helpers import product_schedule, rel_err

CASES = [
    dict(name="ad23_eps_vp", schedule="vp_linear", algo="dpmsolver", order=3, t_end=1e-3, solver_type="dpmsolver"),
    dict(name="ad12_eps_vp", schedule="vp_linear", algo="dpmsolver", order=2, t_end=1e-3, solver_type="dpmsolver"),
    dict(name="ad23_pp_sd", schedule="sd", algo="dpmsolver++", order=3, t_end=None, solver_type="taylor"),
    dict(name="ad12_pp_sd", schedule="sd", algo="dpmsolver++", order=2, t_end=None, solver_type="dpmsolver"),
]


def run(c, device, capsys):
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule(c["schedule"])
    x = seeded((2, 3, 8, 8), 77).to(device)
    s = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type=c["algo"])
    y = s.sample(x, method="adaptive", order=c["order"], t_end=c["t_end"], solver_type=c["solver_type"])
    nfe = int(capsys.readouterr().out.split()[-1])
    return y, nfe


@pytest.fixture(scope="module")
def gold():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "adaptive.npz"))


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_adaptive_host_logic(gold, oracle_backend, capsys, c):
    y, nfe = run(c, "cpu", capsys)
    assert nfe == int(gold[c["name"] + "/nfe"])
    assert rel_err(y.numpy(), gold[c["name"] + "/y"]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_adaptive_gpu(gold, cuda_backend, capsys, c):
    y, nfe = run(c, "cuda:0", capsys)
    assert nfe == int(gold[c["name"] + "/nfe"])
    assert rel_err(y.cpu().numpy(), gold[c["name"] + "/y"]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt", [((5, 3, 64, 64), torch.float32), ((3, 1001), torch.float32), ((4, 4, 64, 64), torch.bfloat16)])
def test_error_norm_kernel(cuda_backend, shape, dt):
    g = torch.Generator().manual_seed(5)
    xh, xl, xp = (torch.randn(shape, generator=g).to(dt) for _ in range(3))
    xl = (xh.float() + 0.01 * xl.float()).to(dt)
    got = float(cuda_backend.error_norm(xh.cuda(), xl.cuda(), xp.cuda(), 0.0078, 0.05).cpu())
    h, l, p = (t.double().numpy() for t in (xh, xl, xp))
    delta = np.maximum(0.0078, 0.05 * np.maximum(np.abs(l), np.abs(p)))
    ref = np.sqrt(np.mean(np.square(((h - l) / delta).reshape(shape[0], -1)), axis=-1)).max()
    assert abs(got - ref) <= 2e-6 * ref


# ---- the controller on the device (csrc/adaptive_ctl.cu) ------------------------------------------------------------
def _adaptive_cfgs(n, seed):
    import random
    from test_random_configs_vs_reference import draw_wide
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        c = draw_wide(rng)
        c.update(method="adaptive", order=rng.choice([2, 3]), atol=rng.choice([0.0078, 0.05]), rtol=rng.choice([0.05, 0.2]),
                 thresholding=False, denoise_to_zero=False)
        if c["schedule"] == "iddpm_cosine":
            continue
        out.append(c)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", range(4))
def test_device_controller_matches_reference(cuda_backend, chunk):
    """Random adaptive configurations (schedules, both algorithms and solver types, all parameterisations, CFG,
    t_start / t_end, batch shapes): the device-side controller takes the accept/reject decisions of the UNMODIFIED
    reference (same NFE) and lands on its sample within the reduction-order / device-libm tolerance."""
    import contextlib
    import io
    import dpm_solver_b200 as new
    from oracle import ref_loader
    from test_random_configs_vs_reference import run_wide
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_loader.load("dpm_solver_pytorch")

    class OnGpu:        # run_wide() builds CPU tensors: move the product arm to the device
        NoiseScheduleVP = new.NoiseScheduleVP

        @staticmethod
        def model_wrapper(net, ns, **kw):
            kw = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
            return new.model_wrapper(net, ns, **kw)

        class DPM_Solver(new.DPM_Solver):
            def sample(self, x, **kw):
                return super().sample(x.cuda(), **kw).cpu()

    seen_device = 0
    for c in _adaptive_cfgs(10, 9000 + chunk):
        out_r, out_n = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(out_r):
            yr, _, _ = run_wide(ref, c)
        if not torch.isfinite(yr).all():
            continue
        before = cuda_backend.launch_count()
        with contextlib.redirect_stdout(out_n):
            yn, _, _ = run_wide(OnGpu, c)
        seen_device += cuda_backend.launch_count() > before
        assert out_n.getvalue().split()[-1] == out_r.getvalue().split()[-1], ("NFE", c)
        assert rel_err(yn.numpy(), yr.numpy()) <= 5e-4, c
    assert seen_device


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_device_controller_syncs_once_per_chunk(gold, cuda_backend, capsys, monkeypatch, c):
    """The only device->host read of the adaptive loop is AdaptiveController.read(): once per `adaptive_chunk`
    iterations. Same NFE and sample as the host controller (the reference's per-iteration decision)."""
    from dpm_solver_b200 import DPM_Solver, model_wrapper, ops
    reads = []
    orig = ops.AdaptiveController.read

    def counting(self):
        r = orig(self)
        reads.append(r)
        return r
    monkeypatch.setattr(ops.AdaptiveController, "read", counting)
    y_dev, nfe_dev = run(c, "cuda:0", capsys)
    iters = reads[-1][2]
    assert len(reads) == -(-iters // DPM_Solver.adaptive_chunk) and reads[-1][0] == 1
    assert nfe_dev == int(gold[c["name"] + "/nfe"]) == iters * c["order"]
    monkeypatch.setattr(DPM_Solver, "adaptive_controller", "host")
    y_host, nfe_host = run(c, "cuda:0", capsys)
    assert nfe_host == nfe_dev
    assert rel_err(y_dev.cpu().numpy(), y_host.cpu().numpy()) <= 1e-5
