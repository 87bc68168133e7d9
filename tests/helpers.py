"""Shared test helpers: run a golden case through the product API or through the oracle."""
import numpy as np
import torch

from cases import SAMPLE_CASES, exact_net, make_betas, seeded, sin_net

CASES = {c["name"]: c for c in SAMPLE_CASES}


def product_schedule(name, dtype=torch.float32):
    from dpm_solver_b200 import NoiseScheduleVP
    kind, betas = make_betas(name)
    if kind == "linear":
        return NoiseScheduleVP("linear", continuous_beta_0=0.1, continuous_beta_1=20.)
    return NoiseScheduleVP("discrete", betas=torch.from_numpy(betas), dtype=dtype)


def oracle_schedule(name, golden=None, xp=None):
    from oracle import dpm_oracle as O
    xp = xp or O.NP
    kind, betas = make_betas(name)
    if kind == "linear":
        return O.VPSchedule("linear", beta_0=0.1, beta_1=20., xp=xp)
    ns = O.VPSchedule.from_betas(betas, xp=xp)
    if golden is not None:  # exact fp32 tables of the reference instance
        ns.set_tables(golden[f"{name}/t_array"], golden[f"{name}/log_alpha_array"])
    return ns


def run_product_case(case, device="cpu", state_dtype=None, model_dtype=None, return_solver=False):
    """sample() of the product on `device` for a golden case -> (y, intermediates, calls)."""
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule(case["schedule"])
    B = case["shape"][0]
    x = seeded(case["shape"], case["seed"]).to(device)
    calls = []
    net0 = sin_net if case["net"] == "sin" else exact_net

    def cast(o):
        return o if model_dtype is None else o.to(model_dtype)

    if case.get("cfg"):
        def net(xx, tt, cc):
            calls.append((float(tt[0]), tuple(xx.shape)))
            return cast(net0(xx.float(), tt) + 0.05 * cc.reshape(-1, 1, 1, 1))
        fn = model_wrapper(net, ns, model_type=case["model_type"], guidance_type="classifier-free",
                           condition=torch.ones(B, 1, device=device),
                           unconditional_condition=torch.zeros(B, 1, device=device), guidance_scale=case["cfg"])
    else:
        def net(xx, tt):
            calls.append((float(tt[0]), tuple(xx.shape)))
            return cast(net0(xx.float(), tt))
        fn = model_wrapper(net, ns, model_type=case["model_type"])
    s = DPM_Solver(fn, ns, algorithm_type=case["algo"],
                   correcting_x0_fn="dynamic_thresholding" if case.get("thresholding") else None,
                   state_dtype=state_dtype)
    y, inter = s.sample(x, steps=case["steps"], order=case["order"], skip_type=case["skip_type"],
                        method=case["method"], lower_order_final=case.get("lower_order_final", True),
                        denoise_to_zero=case.get("denoise_to_zero", False),
                        solver_type=case.get("solver_type", "dpmsolver"), return_intermediate=True,
                        t_end=case.get("t_end"))
    if return_solver:
        return y, inter, calls, s
    return y, inter, calls


def run_oracle_case(case, golden_sched=None, xp=None):
    """The same case through oracle/dpm_oracle.py (numpy) -> (y, intermediates, calls)."""
    from oracle import dpm_oracle as O
    xp = xp or O.NP
    ns = oracle_schedule(case["schedule"], golden_sched, xp)
    x = seeded(case["shape"], case["seed"])
    net0 = sin_net if case["net"] == "sin" else exact_net
    as_t = (lambda a: a) if xp.name == "torch" else (lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    back = (lambda t: t) if xp.name == "torch" else (lambda t: t.numpy())
    B = case["shape"][0]
    if case.get("cfg"):
        cc = torch.cat([torch.zeros(B, 1), torch.ones(B, 1)])
        net = lambda xx, tt: back(net0(as_t(xx), as_t(tt)) + 0.05 * cc.reshape(-1, 1, 1, 1))
    else:
        net = lambda xx, tt: back(net0(as_t(xx), as_t(tt)))
    smp = O.Sampler(ns, net, algorithm_type=case["algo"], model_type=case["model_type"],
                    guidance_scale=case.get("cfg"),
                    thresholding=(0.995, 1.0) if case.get("thresholding") else None)
    x0 = back(x)
    kw = dict(steps=case["steps"], order=case["order"], skip_type=case["skip_type"], t_0=case.get("t_end"),
              solver_type=case.get("solver_type", "dpmsolver"))
    if case["method"] == "multistep":
        y, inter = smp.multistep(x0, lower_order_final=case.get("lower_order_final", True), **kw)
    else:
        y, inter = smp.singlestep(x0, fixed=case["method"] == "singlestep_fixed", **kw)
    if case.get("denoise_to_zero"):
        t0 = 1. / ns.total_N if case.get("t_end") is None else case["t_end"]
        y = O.data_prediction(ns, y, smp.noise(y, xp.asarray([t0])), xp.asarray([t0]), smp.thr)
    return y, inter, smp.calls


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
