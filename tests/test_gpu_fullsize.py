"""Full BASELINE sizes on the GPU, checked through size-independent properties (the numpy oracle
would take minutes there): (1) the fused kernels equal the same chain written as separate torch
CUDA ops (each op rounds once, like the reference's eager chain) bit for bit; (2) the direct and
TMA-ring variants agree bit for bit; (3) dynamic thresholding equals torch.sort-based order
statistics; (4) sample() on a shard equals the matching rows of sample() on the whole batch."""
import numpy as np
import pytest
import torch

from dpm_solver_b200 import ops
from dpm_solver_b200._lib import FORM_DIFF2, FORM_MS3, FORM_NONE
from dpm_solver_b200.ops import StepArgs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CO = dict(a=0.94983894, c0=0.0897649, c1=-0.04488245, c2=0.0021, w0=0.9766731, w1=1.0613433, w2=0.4912, w3=0.4796)


def eager_ms3(x, T0, m1, m2, c):
    D10 = c["w0"] * (T0 - m1)
    D11 = c["w1"] * (m1 - m2)
    dd = D10 - D11
    return ((c["a"] * x + c["c0"] * T0) + c["c1"] * (D10 + c["w2"] * dd)) + c["c2"] * (c["w3"] * dd)


@pytest.mark.parametrize("shape,dt", [((4096, 4, 64, 64), torch.bfloat16), ((4096, 4, 64, 64), torch.float32),
                                      ((1024, 3, 256, 256), torch.float32)])
def test_fused_ms3_step_equals_eager_chain(cuda_backend, shape, dt):
    g = torch.Generator(device=DEV).manual_seed(0)
    mk = lambda: torch.randn(shape, device=DEV, generator=g).to(dt)
    x, ec, eu, m1, m2 = mk(), mk(), mk(), mk(), mk()
    alpha, sigma, scale = 0.83, 0.55, 7.5
    outs = {}
    for variant in (0, 1):
        cuda_backend.set_tuning(variant, 0, 0)
        a = StepArgs(form=FORM_MS3, n_model=2, x=x, xe=x, e_cond=ec, e_uncond=eu, m1=m1, m2=m2, predict_x0=True,
                     guidance=scale, alpha_e=alpha, sigma_e=sigma, want_m_out=True, state_dtype=dt, **CO)
        outs[variant] = cuda_backend.step(a)
    cuda_backend.set_tuning(2, 0, 0)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # eager chain in fp32, one rounding per op; division by a python scalar would become a reciprocal
    # multiply on CUDA, so divide by a 1-element tensor (true division, like the reference's (1,) tensors)
    xf, ecf, euf, m1f, m2f = (t.float() for t in (x, ec, eu, m1, m2))
    eps = euf + scale * (ecf - euf)
    x0 = (xf - sigma * eps) / torch.tensor([alpha], device=DEV)
    T0 = x0.to(dt).float()
    ref = eager_ms3(xf, T0, m1f, m2f, CO).to(dt)
    assert torch.equal(outs[0][0], x0.to(dt))
    assert torch.equal(outs[0][1], ref)
    del outs, ref


def test_full_size_dynamic_threshold(cuda_backend):
    """The C4 quantile at its full size, [1024,3,256,256] in ONE call, bit-exact for all 1024 samples: the two order
    statistics come from torch.sort (in quarters of the batch, sort needs the memory), their interpolation from ATen's
    own CPU lerp -- what torch.quantile computes (:422) -- and a handful of samples go through torch.quantile itself."""
    shape = (1024, 3, 256, 256)
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(shape, device=DEV, generator=g)
    e = torch.randn(shape, device=DEV, generator=g)
    alpha, sigma = 0.37, 0.929
    a = StepArgs(form=FORM_NONE, n_model=1, e_cond=e, xe=x, predict_x0=True, alpha_e=alpha, sigma_e=sigma,
                 per_sample=3 * 256 * 256, state_dtype=torch.float32)
    s = cuda_backend.dynamic_threshold(a, 0.995, 1.0).cpu()
    n = 3 * 256 * 256
    pos = np.float32(0.995) * np.float32(n - 1)
    lo = int(np.floor(pos))
    w = float(np.float32(pos - np.float32(lo)))
    ref = []
    for q in range(4):
        rows = slice(256 * q, 256 * (q + 1))
        x0 = ((x[rows] - sigma * e[rows]) / torch.tensor([alpha], device=DEV)).reshape(256, -1).abs()
        srt = torch.sort(x0, dim=1).values
        vl, vh = srt[:, lo].cpu(), srt[:, lo + 1].cpu()
        ref.append(torch.maximum(torch.lerp(vl, vh, torch.tensor(w)), torch.tensor(1.0)))     # CPU lerp, then :423
        if q == 0:
            tq = torch.quantile(x0[:3].cpu(), 0.995, dim=1).clamp_min(1.0)
            assert torch.equal(s[:3], tq)
        del x0, srt
    assert torch.equal(s, torch.cat(ref))


def test_sample_on_shard_equals_rows_of_full_batch(cuda_backend):
    """T6 on one GPU: rows [lo,hi) of sample(whole batch) == sample(shard), bitwise (no cross-sample coupling)."""
    from cases import make_betas
    from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper
    from dpm_solver_b200.distributed import shard_bounds
    ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("ddpm_linear")[1]))
    B = 64
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(B, 3, 64, 64, device=DEV, generator=g)
    net = lambda xx, tt: 0.1 * xx + ((tt * 0.001) * 0.05 - 0.02).reshape(-1, 1, 1, 1)
    mk = lambda: DPM_Solver(model_wrapper(net, ns), ns, correcting_x0_fn="dynamic_thresholding")
    full = mk().sample(x, steps=12, order=3)
    for r in range(4):
        lo, hi = shard_bounds(B, r, 4)
        part = mk().sample(x[lo:hi].contiguous(), steps=12, order=3)
        assert torch.equal(part, full[lo:hi])


def test_non_default_stream_and_empty(cuda_backend):
    s = torch.cuda.Stream()
    x = torch.randn(8 * 100000, device=DEV)
    m = torch.randn(8 * 100000, device=DEV)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y = ops.lincomb(x, [m], 0.5, [0.25])
    torch.cuda.current_stream().wait_stream(s)
    assert torch.equal(y, 0.5 * x + 0.25 * m)
    e = torch.empty(0, device=DEV)
    assert ops.lincomb(e, [e], 1.0, [1.0]).numel() == 0


def test_beyond_2_31_elements(cuda_backend):
    """Maximum sizes: more than 2^31 elements in one call (64-bit byte offsets, 32-bit packet indices) on both kernel
    variants and the duplicate kernel -- bf16 keeps it at 4.3 GB per tensor. Checked on slices at the start, across
    the 2^31 boundary and at the ragged tail against the unfused fp32 expression rounded to bf16."""
    n = (1 << 31) + 8 * 1000 + 3
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(n, device=DEV, generator=g).bfloat16()
    m0 = torch.randn(n, device=DEV, generator=g).bfloat16()
    a, c0 = 0.9375, -0.40625
    spots = [slice(0, 4096), slice((1 << 31) - 2048, (1 << 31) + 2048), slice(n - 3000, n)]

    def want(sl):
        return (a * x[sl].float() + c0 * m0[sl].float()).bfloat16()

    for variant in (1, 0):
        cuda_backend.set_tuning(variant=variant)
        try:
            out = ops.lincomb(x, [m0], a, [c0])
        finally:
            cuda_backend.set_tuning(variant=2)
        for sl in spots:
            assert torch.equal(out[sl], want(sl)), (variant, sl)
        del out
    del m0
    xx = x.view(1, n)
    dup = cuda_backend.duplicate(xx)
    assert dup.shape == (2, n)
    for sl in spots:
        assert torch.equal(dup[0, sl], x[sl]) and torch.equal(dup[1, sl], x[sl])
