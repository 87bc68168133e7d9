"""Run a BASELINE.json workload (bench.py WORKLOADS: c2 / c3 / c4) through the UNMODIFIED reference and through
the product on the same synthetic inputs.  TEST INFRASTRUCTURE (uses oracle/_ref).

The synthetic network is bench.py's: it returns pre-generated noise banks in rotation, i.e. exact data, so the
reference's fp32 CPU result is reproducible bit for bit by the fp32-state CUDA path. Every op of the path is
element-wise or per-sample, so a batch SLICE of the inputs gives the corresponding slice of the full-batch
output: `slice_rows` picks the rows of a [B,...] (or doubled [2B,...] CFG) bank that belong to samples [0, n)."""
import numpy as np
import torch

from cases import make_betas
from oracle import ref_loader


def schedule(mod, w):
    kind, betas = make_betas(w["schedule"])
    return mod.NoiseScheduleVP("discrete", betas=torch.from_numpy(betas)) if kind == "discrete" else mod.NoiseScheduleVP("linear")


def slice_rows(bank, B, n, cfg):
    """Rows of samples [0, n) in a bank of a B-sample batch (uncond half first under CFG, reference :328)."""
    if not cfg:
        return bank[:n]
    return torch.cat([bank[:n], bank[B:B + n]])


def _solver(mod, w, banks, device, state_dtype=None, **extra):
    ns = schedule(mod, w)
    B = banks[0].shape[0] // (2 if w["cfg"] else 1)
    cnt = [0]
    if w["cfg"]:
        def net(xx, tt, cc):
            cnt[0] += 1
            return banks[cnt[0] % len(banks)]
        fn = mod.model_wrapper(net, ns, guidance_type="classifier-free", condition=torch.ones(B, 1, device=device),
                               unconditional_condition=torch.zeros(B, 1, device=device), guidance_scale=w["cfg"])
    else:
        def net(xx, tt):
            cnt[0] += 1
            return banks[cnt[0] % len(banks)]
        fn = mod.model_wrapper(net, ns)
    kw = dict(algorithm_type=w["algo"], correcting_x0_fn="dynamic_thresholding" if w["thresholding"] else None)
    if state_dtype is not None:
        kw["state_dtype"] = state_dtype
    kw.update(extra)
    return mod.DPM_Solver(fn, ns, **kw), cnt


def sample_kwargs(w):
    return dict(steps=w["steps"], order=w["order"], method=w["method"], skip_type="time_uniform")


def reference_sample(w, x_T, banks, device="cpu"):
    """The unmodified reference, fp32, on `device` ("cpu": the box's host cores)."""
    ref = ref_loader.load("dpm_solver_pytorch")
    x = x_T.detach().to(device=device, dtype=torch.float32)
    bk = [b.detach().to(device=device, dtype=torch.float32) for b in banks]
    s, _ = _solver(ref, w, bk, device)
    return s.sample(x, **sample_kwargs(w))


def product_sample(w, x_T, banks, state_dtype=None):
    import dpm_solver_b200 as new
    s, _ = _solver(new, w, banks, x_T.device, state_dtype=state_dtype)
    return s.sample(x_T, **sample_kwargs(w))


def synthetic(w, B, device, dtype, seed=1234, nbanks=3):
    """bench.py's inputs for a B-sample batch: x_T ~ N(0,1), noise banks ~ N(0,1), rounded to `dtype`."""
    shape = (B,) + tuple(w["shape"][1:])
    g = torch.Generator(device=device).manual_seed(seed)
    x_T = torch.randn(shape, device=device, generator=g).to(dtype)
    nb = 2 if w["cfg"] else 1
    banks = [torch.randn((nb * B,) + shape[1:], device=device, generator=g).to(dtype) for _ in range(nbanks)]
    return x_T, banks


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rms_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))
