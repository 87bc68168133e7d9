"""Reference-rounding mode (dpm_step_desc.raw_round, DPM_Solver(reference_rounding=True)) on the GPU: the
generic kernel's <RND> instantiation against the numpy executor, bit for bit, and one sample() run.

The semantics are pinned on CPU against the unmodified reference
(tests/test_random_configs_vs_reference.py::test_reference_rounding_mode_is_bit_identical); all cases passed on
a B200 in round 1's driver run (39 XPASS), so the tests are strict now."""
import random

import pytest
import torch

from dpm_solver_b200._lib import (FORM_DIFF2, FORM_LIN1, FORM_LIN3, FORM_MS3, FORM_NONE, FORM_SS3T, PARAM_NOISE)
from dpm_solver_b200.ops import StepArgs
from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mdt,code", [(torch.bfloat16, 1), (torch.float16, 2)], ids=["bf16", "f16"])
@pytest.mark.parametrize("form", [FORM_NONE, FORM_LIN1, FORM_LIN3, FORM_DIFF2, FORM_MS3, FORM_SS3T])
@pytest.mark.parametrize("n_model", [0, 1, 2])
def test_raw_round_kernel_equals_executor(cuda_backend, mdt, code, form, n_model):
    if form == FORM_NONE and n_model == 0:
        pytest.skip("NONE needs a network output")
    rng = random.Random(form * 10 + n_model)
    n = 8 * 5000 + 5
    g = torch.Generator().manual_seed(form * 7 + n_model + code)
    f32 = lambda: torch.randn(n, generator=g)
    as_T = lambda: torch.randn(n, generator=g).to(mdt)
    v = [rng.uniform(0.2, 1.5) * rng.choice([-1, 1]) for _ in range(9)]
    rr = code | (4 if form != FORM_NONE else 0)
    a = StepArgs(form=form, n_model=n_model, param=PARAM_NOISE, predict_x0=False, guidance=3.7,
                 a=v[0], c0=v[1], c1=v[2], c2=v[3], w0=v[4], w1=v[5], w2=abs(v[6]), w3=abs(v[7]), w4=abs(v[8]) + 0.1,
                 want_m_out=True, state_dtype=torch.float32, raw_round=rr)
    host = {}
    if form != FORM_NONE:
        host["x"] = f32()
    if n_model == 0:
        host["m0"] = as_T().float()            # buffers hold raw 16-bit outputs, widened
    else:
        host["e_cond"] = as_T()
        if n_model == 2:
            host["e_uncond"] = as_T()
    if form in (FORM_LIN3, FORM_DIFF2, FORM_MS3, FORM_SS3T):
        host["m1"] = as_T().float()
    if form in (FORM_LIN3, FORM_MS3, FORM_SS3T):
        host["m2"] = as_T().float()
    ah = StepArgs(**{**a.__dict__, **host})
    ad = StepArgs(**{**a.__dict__, **{k: t.to(DEV) for k, t in host.items()}})
    mh, oh = OracleBackend().step(ah)
    md, od = cuda_backend.step(ad)
    if mh is not None:
        assert torch.equal(md.cpu(), mh)
    if oh is not None:
        assert torch.equal(od.cpu(), oh)


@pytest.mark.parametrize("algo,method,order,solver_type", [("dpmsolver", "multistep", 3, "dpmsolver"),
                                                          ("dpmsolver", "singlestep", 3, "taylor"),
                                                          ("dpmsolver++", "multistep", 2, "dpmsolver")])
def test_reference_rounding_sample_equals_executor(cuda_backend, oracle_backend_cpu, algo, method, order, solver_type):
    from cases import exact_net, make_betas, seeded
    from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper, ops

    def run(device):
        ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("sd")[1]))

        def net(xx, tt, cc):
            return (exact_net(xx.float(), tt) + 0.05 * cc.reshape(-1, 1, 1, 1)).to(torch.bfloat16)
        fn = model_wrapper(net, ns, guidance_type="classifier-free", condition=torch.ones(3, 1, device=device),
                           unconditional_condition=torch.zeros(3, 1, device=device), guidance_scale=3.7)
        s = DPM_Solver(fn, ns, algorithm_type=algo, reference_rounding=True)
        return s.sample(seeded((3, 4, 16, 16), 11).to(device), steps=9, order=order, method=method, solver_type=solver_type)

    y = run(DEV)
    ops.set_backend(oracle_backend_cpu)
    assert torch.equal(y.cpu(), run("cpu"))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_16bit_x_promotion_mode_equals_executor(cuda_backend, oracle_backend_cpu, dt):
    """x_T handed over in 16 bits with state_dtype=None (staged with the rest of this file): the network sees the
    16-bit tensor at the first evaluation, fp32 afterwards; CUDA path == numpy executor, bit for bit."""
    from cases import exact_net, make_betas, seeded
    from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper, ops

    def run(device):
        ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("sd")[1]))
        seen = []

        def net(xx, tt):
            seen.append(xx.dtype)
            return exact_net(xx.float(), tt).to(xx.dtype)
        s = DPM_Solver(model_wrapper(net, ns), ns, algorithm_type="dpmsolver++")
        y = s.sample(seeded((3, 4, 16, 16), 12).to(dt).to(device), steps=8, order=3)
        return y, seen

    y, seen = run(DEV)
    assert seen[0] == dt and all(d == torch.float32 for d in seen[1:]) and y.dtype == torch.float32
    ops.set_backend(oracle_backend_cpu)
    assert torch.equal(y.cpu(), run("cpu")[0])
