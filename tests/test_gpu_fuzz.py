"""Fuzz: random forms, dtypes, sizes, pointer offsets (misalignment), aliasing and tuning against the
numpy executor, bit for bit. Seeds are fixed, the case list is not hand-picked."""
import random

import numpy as np
import pytest
import torch

from dpm_solver_b200._lib import (FORM_DIFF2, FORM_LIN1, FORM_LIN2, FORM_LIN3, FORM_MS3, FORM_NONE, FORM_SS3T,
                                  PARAM_NOISE, PARAM_SCORE, PARAM_V, PARAM_X_START)
from dpm_solver_b200.ops import StepArgs
from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTS = [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16),
       (torch.float32, torch.bfloat16), (torch.float32, torch.float16), (torch.bfloat16, torch.float16)]


def one_case(rng, be):
    form = rng.choice([FORM_NONE, FORM_LIN1, FORM_LIN2, FORM_LIN3, FORM_DIFF2, FORM_MS3, FORM_SS3T])
    n_model = rng.choice([0, 1, 2]) if form != FORM_NONE else rng.choice([1, 2])
    sdt, mdt = rng.choice(DTS)
    n = rng.choice([rng.randint(1, 64), rng.randint(65, 5000), 8 * rng.randint(100, 40000) + rng.randint(0, 7),
                    8 * 148 * 1024 + rng.randint(0, 4096)])
    off = rng.choice([0, 0, 0, 1, 3, 4, 8])                # element offset into a larger allocation
    g = torch.Generator().manual_seed(rng.randint(0, 1 << 30))
    mk = lambda dt: (torch.randn(n + 16, generator=g) * rng.choice([0.1, 1.0, 30.0])).to(dt)
    param = rng.choice([PARAM_NOISE, PARAM_NOISE, PARAM_X_START, PARAM_V, PARAM_SCORE])
    px0 = rng.random() < 0.6
    v = [rng.uniform(0.2, 1.5) * rng.choice([-1, 1]) for _ in range(9)]
    a = StepArgs(form=form, n_model=n_model, param=param, predict_x0=px0 and n_model > 0, c0_on_old=rng.random() < 0.5,
                 guidance=rng.choice([1.0, 3.5, 7.5]), alpha_e=rng.uniform(0.004, 1.0), sigma_e=rng.uniform(0.03, 1.0),
                 a=v[0], c0=v[1], c1=v[2], c2=v[3], w0=v[4], w1=v[5], w2=abs(v[6]), w3=abs(v[7]), w4=abs(v[8]) + 0.1,
                 want_m_out=rng.random() < 0.7, state_dtype=sdt)
    host, devt = {}, {}

    def put(name, dt):
        t = mk(dt)
        host[name] = t[off:off + n]
        devt[name] = t.to(DEV)[off:off + n]

    if form != FORM_NONE:
        put("x", sdt)
    if n_model == 0:
        put("m0", sdt)
    else:
        put("e_cond", mdt)
        if n_model == 2:
            put("e_uncond", mdt)
        if a.predict_x0 or param in (PARAM_X_START, PARAM_V):
            if form == FORM_NONE or rng.random() < 0.4:
                put("xe", sdt)
            else:
                host["xe"], devt["xe"] = host["x"], devt["x"]
    if form in (FORM_LIN2, FORM_LIN3, FORM_DIFF2, FORM_MS3, FORM_SS3T):
        put("m1", sdt)
    if form in (FORM_LIN3, FORM_MS3, FORM_SS3T):
        put("m2", sdt)
    ah, ad = StepArgs(**{**a.__dict__, **host}), StepArgs(**{**a.__dict__, **devt})
    if a.predict_x0 and n_model > 0 and rng.random() < 0.3 and n >= 64:
        ps = rng.choice([d for d in (8, 16, 24, 1, 7) if n % d == 0] or [n])
        thr = torch.rand(n // ps, generator=g) * 2 + 0.3
        ah.thr, ah.per_sample, ad.thr, ad.per_sample = thr, ps, thr.to(DEV), ps
    be.set_tuning(rng.choice([0, 1, 2]), rng.choice([0, 64, 128, 256, 512]), rng.choice([0, 1, 2, 4]))
    ref_m, ref_o = OracleBackend().step(ah)
    got_m, got_o = be.step(ad)
    for r, q in ((ref_m, got_m), (ref_o, got_o)):
        assert (r is None) == (q is None)
        if r is not None:
            w = torch.int16 if r.element_size() == 2 else torch.int32
            assert torch.equal(q.cpu().view(w), r.view(w)), (form, n_model, sdt, mdt, n, off, param, px0)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_step_kernels(cuda_backend, seed):
    rng = random.Random(4242 + seed)
    try:
        for _ in range(40):
            one_case(rng, cuda_backend)
    finally:
        cuda_backend.set_tuning(2, 0, 0)
