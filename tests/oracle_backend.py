"""A numpy executor for `StepArgs` -- TEST INFRASTRUCTURE.

Restates, element-wise in IEEE fp32 (numpy never fuses a*b+c), the step semantics documented in
include/dpm_solver_b200.h (`dpm_form`, `dpm_param`, `dpm_step_desc`). Installed with
`dpm_solver_b200.ops.set_backend()` it lets the CPU test-suite drive the product's host logic
(plan, step ordering, buffer rotation, hooks) end to end without a GPU, and on the GPU box it is
the per-launch checker the CUDA kernels are compared against, bit for bit.

The quantile comes from oracle/dpm_oracle.py (sort based, torch.quantile semantics).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import dpm_oracle as O  # noqa: E402
from dpm_solver_b200._lib import (FORM_DIFF2, FORM_LIN1, FORM_LIN2, FORM_LIN3, FORM_MS3, FORM_NONE,  # noqa: E402
                                  FORM_SS3T, PARAM_SCORE, PARAM_V, PARAM_X_START)

f32 = np.float32
_RR_DTYPE = {1: torch.bfloat16, 2: torch.float16}   # dpm_dtype codes carried by raw_round


def _np(t):
    return None if t is None else t.detach().float().cpu().numpy()


def _round(v, dtype):
    """fp32 values as they read back from `dtype` storage."""
    t = torch.from_numpy(np.ascontiguousarray(v))
    return t.to(dtype).float().numpy() if dtype != torch.float32 else v


class OracleBackend:
    name = "numpy-oracle"

    def __init__(self):
        self.launches = 0
        self.log = []   # (form, n_model) per launch

    def launch_count(self):
        return self.launches

    def set_tuning(self, *a, **k):
        pass

    # ---- model value (dpm_step_desc, n_model >= 1) ----
    @staticmethod
    def _convert(a, out, xe):
        al, sg = f32(a.alpha_e), f32(a.sigma_e)
        if a.param == PARAM_X_START:
            return (xe - al * out) / sg
        if a.param == PARAM_V:
            return al * out + sg * xe
        if a.param == PARAM_SCORE:
            return (-sg) * out
        return out

    def _model_value(self, a, thr=None):
        xe = _np(a.xe if a.xe is not None else a.x)
        ec, eu = _np(a.e_cond), _np(a.e_uncond)
        eps = self._convert(a, ec, xe)
        if a.n_model == 2:
            epu = self._convert(a, eu, xe)
            rr = getattr(a, "raw_round", 0)
            if rr and a.param not in (PARAM_X_START, PARAM_V, PARAM_SCORE):
                # reference-rounding mode (dpm_step_desc.raw_round): three 16-bit ops
                dt = _RR_DTYPE[rr & 3]
                d = _round((eps - epu).astype(f32), dt)
                eps = _round((epu + _round((f32(a.guidance) * d).astype(f32), dt)).astype(f32), dt)
            else:
                eps = epu + f32(a.guidance) * (eps - epu)
        if a.predict_x0:
            x0 = (xe - f32(a.sigma_e) * eps) / f32(a.alpha_e)
            if thr is not None:
                s = np.repeat(thr.reshape(-1), a.per_sample).reshape(x0.shape)
                x0 = np.minimum(np.maximum(x0, -s), s) / s
            return x0.astype(f32)
        return eps.astype(f32)

    # ---- API of CudaBackend ----
    def step(self, a):
        self.launches += 1
        self.log.append((a.form, a.n_model))
        ref = a.reference_tensor()
        sdt = a.state_dtype
        if sdt is None:
            st = a.state_tensors()
            sdt = st[0].dtype if st else a.e_cond.dtype
        if a.n_model > 0:
            thr = _np(a.thr) if a.thr is not None else None
            T0 = _round(self._model_value(a, thr), sdt)
        else:
            T0 = _np(a.m0)
        m_out = out = None
        if a.n_model > 0 and (a.want_m_out or a.form == FORM_NONE):
            m_out = torch.from_numpy(T0.copy()).to(sdt).reshape(ref.shape)
        if a.form != FORM_NONE:
            x, m1, m2 = _np(a.x), _np(a.m1), _np(a.m2)
            A, c0, c1, c2 = f32(a.a), f32(a.c0), f32(a.c1), f32(a.c2)
            w0, w1, w2, w3, w4 = (f32(v) for v in (a.w0, a.w1, a.w2, a.w3, a.w4))
            rr = getattr(a, "raw_round", 0)
            diff = (lambda u, v: _round((u - v).astype(f32), _RR_DTYPE[rr & 3])) if rr & 4 else (lambda u, v: u - v)
            if a.form == FORM_LIN1:
                o = A * x + c0 * T0
            elif a.form == FORM_LIN2:
                o = (A * x + c0 * T0) + c1 * m1
            elif a.form == FORM_LIN3:
                o = ((A * x + c0 * T0) + c1 * m1) + c2 * m2
            elif a.form == FORM_DIFF2:
                D = w0 * diff(T0, m1)
                o = (A * x + c0 * (m1 if a.c0_on_old else T0)) + c1 * D
            elif a.form == FORM_MS3:
                D10 = w0 * diff(T0, m1)
                D11 = w1 * diff(m1, m2)
                dd = D10 - D11
                o = ((A * x + c0 * T0) + c1 * (D10 + w2 * dd)) + c2 * (w3 * dd)
            elif a.form == FORM_SS3T:
                if rr & 4:      # python-float / 0-dim r1, r2 do not promote: D1, D2 entirely in the 16-bit type
                    R = lambda v: _round(np.asarray(v, dtype=f32), _RR_DTYPE[rr & 3])
                    D10, D11 = R(w0 * R(m1 - m2)), R(w1 * R(T0 - m2))
                    D1 = R(R(R(w2 * D10) - R(w3 * D11)) / w4)
                    D2 = R(R(f32(2) * R(D11 - D10)) / w4)
                else:
                    D10 = w0 * (m1 - m2)
                    D11 = w1 * (T0 - m2)
                    D1 = (w2 * D10 - w3 * D11) / w4
                    D2 = (f32(2) * (D11 - D10)) / w4
                o = ((A * x + c0 * m2) + c1 * D1) + c2 * D2
            else:
                raise ValueError(a.form)
            out = torch.from_numpy(np.ascontiguousarray(o.astype(f32))).to(sdt).reshape(ref.shape)
            if a.out is not None:
                a.out.copy_(out.reshape(a.out.shape))
                out = a.out
            if a.out2 is not None:
                a.out2.copy_(out.reshape(a.out2.shape))
        return m_out, out

    def diffedit_corrector(self, x, x0, mask, alpha, sigma, generator=None):
        """x*mask + (1 - mask)*(alpha*x0 + sigma*randn_like(x0)) in numpy fp32, the noise from torch's CPU generator
        exactly as `stochastic_encode` draws it (sampler.py:92-96 -> add_noise :1023-1024)."""
        self.launches += 1
        noise = torch.randn((1, *x0.shape), generator=generator)[0].numpy()
        xf, x0f, m = _np(x), _np(x0), _np(mask)
        inter = f32(alpha) * x0f + f32(sigma) * noise
        out = xf * m + (f32(1) - m) * inter
        return torch.from_numpy(np.ascontiguousarray(out.astype(f32))).to(x.dtype)

    def duplicate(self, x):
        self.launches += 1
        return torch.cat([x] * 2)

    def error_norm(self, x_higher, x_lower, x_prev, atol, rtol):
        """dpm_solver_adaptive :999-1001 in numpy fp32."""
        self.launches += 1
        xh, xl, xp = _np(x_higher), _np(x_lower), _np(x_prev)
        delta = np.maximum(f32(atol), f32(rtol) * np.maximum(np.abs(xl), np.abs(xp)))
        v = ((xh - xl) / delta).reshape(xh.shape[0], -1)
        e = np.sqrt(np.mean(np.square(v), axis=-1, dtype=np.float32))
        return torch.tensor([float(e.max())], dtype=torch.float32)

    def dynamic_threshold(self, a, q, max_val):
        self.launches += 1
        self.log.append(("quantile", a.n_model))
        x0 = self._model_value(a, None)
        B = x0.size // a.per_sample
        s = np.maximum(O.quantile_abs(x0.reshape(B, -1), q), f32(max_val))
        return torch.from_numpy(s.astype(f32))
