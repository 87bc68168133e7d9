"""VP noise schedule scalars: log(alpha_t), alpha_t, sigma_t, lambda_t and lambda^-1.

API-compatible with `NoiseScheduleVP` of the reference (dpm_solver_pytorch.py:6-167) -- same
constructor, attributes and methods -- but every method is a handful of vectorised ops: the
reference's `interpolate_fn` (cat + sort + argmin + 3 where + 4 gather on a [N,1,K+1] tensor,
:1253-1292) becomes one `searchsorted` + four gathers. The arithmetic keeps the reference's fp32
operation order (`y0 + (x - x0) * (y1 - y0) / (x1 - x0)`, :1291), so results are bit-identical;
the solver evaluates these on HOST tensors once per run (see plan.py), never per step on device.
"""
from __future__ import annotations

import torch

__all__ = ["NoiseScheduleVP", "interpolate_fn", "expand_dims"]


def _piecewise_linear(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """y(x) for ascending keypoints xp/yp (1-D), linear extrapolation outside (reference :1253).

    The reference finds the bracket by sorting [x, xp...] (x first, so on ties x precedes the
    equal keypoint): with i = #{xp < x} the bracket is [xp[i-1], xp[i]], clamped to the first /
    last interval. searchsorted(right=False) returns exactly that i.
    """
    K = xp.shape[0]
    i = torch.searchsorted(xp, x.contiguous(), right=False)
    j0 = (i - 1).clamp_(0, K - 2)
    x0, x1 = xp[j0], xp[j0 + 1]
    y0, y1 = yp[j0], yp[j0 + 1]
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0)


def interpolate_fn(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Drop-in for the reference utility (:1253-1292): x [N,C], xp/yp [C,K] -> [N,C]."""
    cols = [_piecewise_linear(x[:, c], xp[c], yp[c]) for c in range(xp.shape[0])]
    return torch.stack(cols, dim=1)


def expand_dims(v: torch.Tensor, dims: int) -> torch.Tensor:
    """[N] -> [N,1,...,1] with `dims` dimensions (reference :1295-1305)."""
    return v[(...,) + (None,) * (dims - 1)]


class NoiseScheduleVP:
    """Forward VP SDE wrapper; see the reference docstring (:16-92) for the maths.

    schedule='discrete': piecewise-linear log(alpha) over t_i = (i+1)/N from `betas` or
    `alphas_cumprod` (tail clipped where lambda < -5.1, :114-125); schedule='linear': the
    continuous VPSDE with beta_0, beta_1 (:134); schedule='cosine' is accepted for the older
    vendored copies of the solver (examples/stable-diffusion/.../dpm_solver.py:114-175).
    """

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20., dtype=torch.float32):
        if schedule not in ["discrete", "linear", "cosine"]:
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'discrete' or "
                             "'linear'".format(schedule))
        self.schedule = schedule
        self.T = 1.
        self._tables = {}
        if schedule == "discrete":
            if betas is not None:
                log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(alphas_cumprod)
            log_alphas = self.numerical_clip_alpha(log_alphas)
            self.log_alpha_array = log_alphas.reshape((1, -1,)).to(dtype=dtype)
            self.total_N = self.log_alpha_array.shape[1]
            self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:].reshape((1, -1)).to(dtype=dtype)
        else:
            self.total_N = 1000
            self.beta_0 = continuous_beta_0
            self.beta_1 = continuous_beta_1
            if schedule == "cosine":
                import math
                self.cosine_s = 0.008
                self.cosine_beta_max = 999.
                self.cosine_t_max = math.atan(self.cosine_beta_max * (1. + self.cosine_s) / math.pi) * 2. \
                    * (1. + self.cosine_s) / math.pi - self.cosine_s
                self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1. + self.cosine_s) * math.pi / 2.))
                self.T = 0.9946

    # -- construction helper (:114-125) ----------------------------------------------------
    def numerical_clip_alpha(self, log_alphas, clipped_lambda=-5.1):
        """Drop the tail of the table where lambda_t < clipped_lambda (cosine-style schedules)."""
        log_sigmas = 0.5 * torch.log(1. - torch.exp(2. * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), clipped_lambda))
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        return log_alphas

    def _table(self, device):
        """(t, log_alpha, flipped log_alpha, flipped t) as 1-D tensors on `device` (cached)."""
        key = str(device)
        tab = self._tables.get(key)
        ver = (self.log_alpha_array._version, self.t_array._version)
        if tab is None or tab[4] is not self.log_alpha_array or tab[5] is not self.t_array or tab[6] != ver:
            t = self.t_array.to(device).reshape(-1)
            la = self.log_alpha_array.to(device).reshape(-1)
            tab = (t.contiguous(), la.contiguous(), torch.flip(la, [0]).contiguous(),
                   torch.flip(t, [0]).contiguous(), self.log_alpha_array, self.t_array, ver)
            self._tables[key] = tab
        return tab

    # -- scalar functions (:127-167) ---------------------------------------------------------
    def marginal_log_mean_coeff(self, t):
        """log(alpha_t) for continuous-time labels t (any shape; flattened like the reference)."""
        if self.schedule == "discrete":
            tt, la = self._table(t.device)[:2]
            return _piecewise_linear(t.reshape(-1), tt, la)
        elif self.schedule == "linear":
            return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        else:
            import math
            log_alpha_fn = lambda s: torch.log(torch.cos((s + self.cosine_s) / (1. + self.cosine_s) * math.pi / 2.))
            return log_alpha_fn(t) - self.cosine_log_alpha_0

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        log_mean_coeff = self.marginal_log_mean_coeff(t)
        log_std = 0.5 * torch.log(1. - torch.exp(2. * log_mean_coeff))
        return log_mean_coeff - log_std

    def inverse_lambda(self, lamb):
        """t(lambda): closed form for 'linear' (:161-163), table inversion for 'discrete' (:165-166)."""
        if self.schedule == "linear":
            tmp = 2. * (self.beta_1 - self.beta_0) * torch.logaddexp(-2. * lamb, torch.zeros((1,)).to(lamb))
            Delta = self.beta_0 ** 2 + tmp
            return tmp / (torch.sqrt(Delta) + self.beta_0) / (self.beta_1 - self.beta_0)
        elif self.schedule == "discrete":
            log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)).to(lamb.device), -2. * lamb)
            la_f, t_f = self._table(lamb.device)[2:4]
            return _piecewise_linear(log_alpha.reshape(-1), la_f, t_f)
        else:
            import math
            log_alpha = -0.5 * torch.logaddexp(-2. * lamb, torch.zeros((1,)).to(lamb))
            t_fn = lambda la: torch.arccos(torch.exp(la + self.cosine_log_alpha_0)) * 2. \
                * (1. + self.cosine_s) / math.pi - self.cosine_s
            return t_fn(log_alpha)
