"""Shared definitions of the golden cases: schedules, deterministic networks, sample() configs.

Imported by make_golden.py (which runs them through the reference) and by the tests (which run
the same cases through the oracle and through the CUDA path)."""
import math

import numpy as np
import torch

SCHEDULES = ("sd", "ddpm_linear", "iddpm_cosine", "vp_linear")
UPDATE_SHAPE = (2, 4, 8, 8)


def make_betas(name):
    """('discrete', float64 betas) or ('linear', None)."""
    if name == "sd":        # Stable-Diffusion v1 scaled-linear
        return "discrete", np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    if name == "ddpm_linear":
        return "discrete", np.linspace(1e-4, 2e-2, 1000, dtype=np.float64)
    if name == "iddpm_cosine":  # improved-DDPM cosine, betas clipped at 0.999 -> numerical_clip_alpha trims the tail
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return "discrete", np.asarray([min(1 - f((i + 1) / 1000) / f(i / 1000), 0.999) for i in range(1000)],
                                      dtype=np.float64)
    if name == "vp_linear":
        return "linear", None
    raise KeyError(name)


def seeded(shape, seed):
    if isinstance(shape, int):
        shape = (shape,)
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def sin_net(x, t_in):
    """The probe network of SURVEY.md section 8c."""
    return 0.1 * x + 0.01 * torch.sin(t_in).reshape(-1, 1, 1, 1)


def exact_net(x, t_in):
    """Only IEEE mul/add: bit-identical on CPU and CUDA, so loop parity can be checked exactly."""
    return 0.1 * x + ((t_in * 0.001) * 0.05 - 0.02).reshape(-1, 1, 1, 1)


def _c(name, **kw):
    d = dict(name=name, schedule="sd", shape=(2, 4, 16, 16), seed=1234, net="exact", algo="dpmsolver++",
             method="multistep", order=2, steps=20, skip_type="time_uniform", model_type="noise")
    d.update(kw)
    return d


SAMPLE_CASES = [
    # BASELINE config 1 and the known answers of SURVEY.md section 8c
    _c("c1_pp2m_sin", shape=(8, 4, 64, 64), net="sin"),
    _c("pp3m_sin", net="sin", order=3),
    _c("eps3s_sin", net="sin", algo="dpmsolver", method="singlestep", order=3, steps=15),
    _c("eps2m_sin", net="sin", algo="dpmsolver"),
    # multistep
    _c("pp2m", traj=True), _c("pp3m", order=3, traj=True), _c("pp1m", order=1, steps=10),
    _c("pp3m_8", order=3, steps=8), _c("pp3m_8_nolof", order=3, steps=8, lower_order_final=False),
    _c("pp2m_3", order=2, steps=3),
    _c("eps2m", algo="dpmsolver"), _c("eps3m", algo="dpmsolver", order=3),
    _c("pp2m_taylor", solver_type="taylor"), _c("eps2m_taylor", algo="dpmsolver", solver_type="taylor"),
    _c("pp2m_logsnr", skip_type="logSNR", steps=15), _c("pp2m_quad", skip_type="time_quadratic", steps=10),
    _c("pp2m_ddpm", schedule="ddpm_linear"), _c("pp2m_cosine", schedule="iddpm_cosine"),
    _c("pp2m_vp", schedule="vp_linear", t_end=1e-3), _c("pp2m_d2z", denoise_to_zero=True),
    # singlestep
    _c("eps3s", algo="dpmsolver", method="singlestep", order=3, steps=15, traj=True),
    _c("eps3s_vp_logsnr", algo="dpmsolver", method="singlestep", order=3, steps=15, schedule="vp_linear",
       skip_type="logSNR", t_end=1e-3),
    _c("pp3s_taylor", method="singlestep", order=3, steps=20, solver_type="taylor"),
    _c("eps3s_taylor", algo="dpmsolver", method="singlestep", order=3, steps=9, solver_type="taylor"),
    _c("pp2s_7", method="singlestep", order=2, steps=7), _c("pp3s_6", method="singlestep", order=3, steps=6),
    _c("eps2s_taylor", algo="dpmsolver", method="singlestep", order=2, steps=8, solver_type="taylor"),
    _c("eps2_fixed", algo="dpmsolver", method="singlestep_fixed", order=2, steps=10),
    _c("pp1s", method="singlestep", order=1, steps=5),
    # classifier-free guidance (config 3: DPM-Solver-3 singlestep, scale 7.5)
    _c("eps3s_cfg", algo="dpmsolver", method="singlestep", order=3, steps=15, cfg=7.5, traj=True),
    _c("pp2m_cfg", cfg=7.5), _c("pp2m_cfg_v", cfg=3.0, model_type="v"),
    # parameterisations
    _c("pp2m_v", model_type="v"), _c("pp2m_xstart", model_type="x_start"), _c("eps2m_score", algo="dpmsolver", model_type="score"),
    # dynamic thresholding (config 4: ++3M, pixel space, DDPM linear schedule)
    _c("pp3m_thr", schedule="ddpm_linear", shape=(2, 3, 16, 16), order=3, thresholding=True, traj=True),
    _c("pp2s_thr", schedule="ddpm_linear", shape=(2, 3, 16, 16), method="singlestep", order=2, steps=10, thresholding=True),
]
