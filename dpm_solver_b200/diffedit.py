"""Fused DiffEdit corrector: a `correcting_xt_fn` for `DPM_Solver` / `DPMSolverSampler.sample`.

The reference's inpainting example (examples/stable-diffusion/scripts/diffedit_inpaint.ipynb, `corrector_fn`)
re-noises the encoded source image at every solver step and blends it in outside the edit mask:

    def corrector_fn(x, t, step):
        ratio = sampler.time_to_ratio(t)
        stochastic_intermediate = sampler.stochastic_encode(init_latent, ratio)   # alpha*x0 + sigma*randn (sampler.py:92-96)
        return x * mask + (1 - mask) * stochastic_intermediate

-- a randn, the three ops of add_noise and four blend ops, eight full-tensor passes. `DiffEditCorrector` is the
same function as ONE kernel (csrc/philox.cu): the noise is drawn in registers by the generator torch.randn uses,
so for the same generator state the result is bit-identical to the expression above.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["DiffEditCorrector"]


class DiffEditCorrector:
    """`DiffEditCorrector(noise_schedule, init_latent, mask, time_fn)(x, t, step)`.

    time_fn maps the solver's current time `t` (a tensor) to the time whose noise level the source image is
    re-noised to -- the notebook's `lambda t: sampler.ratio_to_time(sampler.time_to_ratio(t))`; default: `t` itself.
    The (alpha, sigma) of each solver step are cached after the first run, so steady-state calls read nothing
    back from the device."""

    def __init__(self, noise_schedule, init_latent, mask, time_fn=None, generator=None):
        self.noise_schedule = noise_schedule
        self.init_latent = init_latent
        self.mask = mask.to(torch.float32).contiguous()
        self.time_fn = time_fn
        self.generator = generator
        self._scalars = {}

    def _alpha_sigma(self, t, step):
        hit = self._scalars.get(step)
        if hit is None:
            te = t if self.time_fn is None else self.time_fn(t)
            te = torch.as_tensor(te, dtype=torch.float32).reshape(-1)[:1].cpu()      # one read-back per distinct step
            # stochastic_encode rebuilds the label in the latent's dtype (sampler.py:94) before add_noise
            te = te.to(self.init_latent.dtype).to(torch.float32) if self.init_latent.dtype != torch.float32 else te
            ns = self.noise_schedule
            hit = (float(ns.marginal_alpha(te)), float(ns.marginal_std(te)), float(te))
            self._scalars[step] = hit
        return hit

    def __call__(self, x, t, step):
        alpha, sigma, _ = self._alpha_sigma(t, step)
        be = ops.backend()
        x0 = self.init_latent
        if x0.dtype != x.dtype:
            x0 = x0.to(x.dtype)
        return be.diffedit_corrector(x, x0.expand(x.shape) if x0.shape != x.shape else x0, self.mask.to(x.device),
                                     alpha, sigma, self.generator)
