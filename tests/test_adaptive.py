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
