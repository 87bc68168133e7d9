"""The reference's three example adapters, loaded UNMODIFIED from oracle/_ref with their model zoos stubbed
(SURVEY 8b "who calls it" -- the drop-in acceptance targets):

  * Stable Diffusion  `DPMSolverSampler`      examples/stable-diffusion/ldm/models/diffusion/dpm_solver/sampler.py
  * score_sde         `get_dpm_solver_sampler` examples/score_sde_pytorch/sampling.py:505-558
  * guided-diffusion  `Diffusion.sample_image` examples/ddpm_and_guided-diffusion/runners/diffusion.py:524-639

Each loader takes the solver module the adapter should run on: the reference's own copy or dpm_solver_b200.
"""
import contextlib
import importlib.util
import sys
import types
import warnings

import torch

from oracle import ref_loader


def available():
    return all(ref_loader.available(n) for n in ("sd_sampler", "sd_dpm_solver", "score_sde_sampling", "guided_runner",
                                                 "guided_sampler", "dpm_solver_pytorch"))


@contextlib.contextmanager
def _stubbed(mods):
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _exec(spec):
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def reference_solver(which="root"):
    """root: dpm_solver_pytorch.py; sd: the older copy vendored next to the SD adapter; guided: the copy the
    guided-diffusion runner imports."""
    name = {"root": "dpm_solver_pytorch", "sd": "sd_dpm_solver", "guided": "guided_sampler"}[which]
    return ref_loader.load(name, fresh=True)


# ---- Stable Diffusion ------------------------------------------------------------------------------
def load_sd_adapter(solver_module, tag, device):
    """sampler.py inside a synthetic package whose `.dpm_solver` is `solver_module`."""
    pkg_name = "_sd_adapter_" + tag
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = []
    sys.modules[pkg_name] = pkg
    sys.modules[pkg_name + ".dpm_solver"] = solver_module
    mod = _exec(ref_loader.spec("sd_sampler", pkg_name + ".sampler"))
    if torch.device(device).type != "cuda":
        # the adapter pins its buffers to "cuda" (sampler.py:23-27); the CPU acceptance run keeps them where they are
        mod.DPMSolverSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    return mod


class StubLatentDiffusion:
    """What DPMSolverSampler touches: alphas_cumprod, betas.device, device, apply_model."""

    def __init__(self, device="cpu"):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
        self.betas = betas.float().to(device)
        self.alphas_cumprod = torch.cumprod(1 - betas, 0).float().to(device)
        self.device = torch.device(device)
        self.calls = []

    def apply_model(self, x, t, c):
        self.calls.append((float(t[0]), tuple(x.shape)))
        return 0.1 * x + ((t * 0.001) * 0.05 - 0.02).reshape(-1, 1, 1, 1) + 0.05 * c.reshape(-1, 1, 1, 1)


# ---- score_sde ---------------------------------------------------------------------------------------
def load_score_sde_sampling(solver_module, tag):
    mutils = types.ModuleType("models.utils")
    mutils.from_flattened_numpy = mutils.to_flattened_numpy = mutils.get_score_fn = lambda *a, **k: None
    mutils.get_noise_fn = lambda sde, model, train=False, continuous=True: (lambda x, t: model(x, t))
    models = types.ModuleType("models")
    models.utils = mutils
    with _stubbed({"models": models, "models.utils": mutils, "sde_lib": types.ModuleType("sde_lib"),
                   "dpm_solver": solver_module}):
        return _exec(ref_loader.spec("score_sde_sampling", "_score_sde_sampling_" + tag))


class StubVPSDE:
    beta_0, beta_1, T = 0.1, 20.0, 1.0

    def prior_sampling(self, shape):
        return torch.randn(*shape, generator=torch.Generator().manual_seed(21))


# ---- guided-diffusion runner -------------------------------------------------------------------------
class _Anything(types.ModuleType):
    """Stub module: every attribute is a harmless placeholder (the runner imports a model zoo, datasets,
    FID code, blobfile, tkinter, torchvision at the top; sample_image uses none of them)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def load_guided_runner(solver_module, tag):
    names = ["tkinter", "blobfile", "tqdm", "models", "models.diffusion", "models.improved_ddpm", "models.improved_ddpm.unet",
             "models.guided_diffusion", "models.guided_diffusion.unet", "models.ema", "functions", "functions.losses",
             "functions.ckpt_util", "functions.denoising", "datasets", "evaluate", "evaluate.fid_score", "torchvision",
             "torchvision.utils"]
    stubs = {n: _Anything(n) for n in names}
    stubs["functions"].get_optimizer = lambda *a, **k: None
    pkg = types.ModuleType("dpm_solver")
    pkg.__path__ = []
    pkg.sampler = solver_module
    stubs["dpm_solver"] = pkg
    stubs["dpm_solver.sampler"] = solver_module       # `from dpm_solver.sampler import ...` runs inside sample_image
    with _stubbed(stubs):
        mod = _exec(ref_loader.spec("guided_runner", "_guided_runner_" + tag))

    def sample_image(self_stub, *a, **k):
        with _stubbed(stubs):
            return mod.Diffusion.sample_image(self_stub, *a, **k)
    return mod, sample_image


def guided_self(betas, sample_type="dpmsolver++", thresholding=True, timesteps=12, order=3, method="multistep",
                skip_type="time_uniform", cond_class=True, scale=2.0, denoise=False, solver_type="dpmsolver",
                lower_order_final=True, fixed_class=3):
    """The attributes of the runner object that sample_image reads (diffusion.py:524-639)."""
    ns = types.SimpleNamespace
    args = ns(skip=1, scale=scale, fixed_class=fixed_class, sample_type=sample_type, skip_type=skip_type, timesteps=timesteps,
              thresholding=thresholding, denoise=denoise, dpm_solver_order=order, dpm_solver_method=method,
              lower_order_final=lower_order_final, dpm_solver_type=solver_type, dpm_solver_atol=0.0078,
              dpm_solver_rtol=0.05, eta=0.0)
    config = ns(sampling=ns(classifier_scale=1.0, cond_class=cond_class), data=ns(num_classes=10), model=ns(out_channels=6))
    return ns(args=args, config=config, betas=betas, num_timesteps=betas.shape[0])


def guided_net(x, t, y=None):
    """6-channel output (mean, variance) like improved-DDPM / guided-diffusion; exact IEEE ops only."""
    out = 0.1 * x + ((t * 0.001) * 0.05 - 0.02).reshape(-1, 1, 1, 1)
    if y is not None:
        out = out + (y.to(x.dtype) * 0.015625).reshape(-1, 1, 1, 1)
    return torch.cat([out, out * 0.5], dim=1)


def guided_classifier(x, t):
    """A smooth differentiable classifier with 10 classes built from exact-ish ops; its gradient runs through
    autograd on whichever device x lives on."""
    feat = x.mean(dim=(2, 3))                              # [B, C]
    w = torch.linspace(-1.0, 1.0, 10 * feat.shape[1], dtype=x.dtype, device=x.device).reshape(feat.shape[1], 10)
    return feat @ w + (t * 0.001).reshape(-1, 1)
