"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference (LuChengTHU/dpm-solver, dpm_solver_pytorch.py) ships no tests or golden vectors, so
parity is pinned on outputs of the reference itself, evaluated on CPU in fp32 with the
container's torch. The fixtures are small .npz files plus cases.json describing them; the GPU box
has no /root/reference and only ever reads these files.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("DPM_REFERENCE", "/root/reference"))
from dpm_solver_pytorch import DPM_Solver, NoiseScheduleVP, model_wrapper  # noqa: E402

sys.path.insert(0, HERE)
from cases import (SAMPLE_CASES, SCHEDULES, UPDATE_SHAPE, exact_net, make_betas, seeded,  # noqa: E402
                   sin_net)

torch.set_num_threads(1)


def np32(t):
    return t.detach().cpu().numpy()


def build_schedule(name):
    kind, betas = make_betas(name)
    if kind == "linear":
        return NoiseScheduleVP("linear", continuous_beta_0=0.1, continuous_beta_1=20.)
    return NoiseScheduleVP("discrete", betas=torch.from_numpy(betas))


def gen_schedules():
    out = {}
    for name in SCHEDULES:
        ns = build_schedule(name)
        if ns.schedule == "discrete":
            out[f"{name}/t_array"] = np32(ns.t_array.reshape(-1))
            out[f"{name}/log_alpha_array"] = np32(ns.log_alpha_array.reshape(-1))
            ties = ns.t_array.reshape(-1)[[0, 1, 17, ns.total_N // 2, ns.total_N - 2, ns.total_N - 1]]
        else:
            ties = torch.tensor([1e-3, 0.5, 1.0])
        q = torch.cat([torch.linspace(1., 1e-3, 21), ties, torch.tensor([1.2, 5e-4, 1e-5, 0.9946]),
                       seeded(64, 7).abs().clamp(1e-4, 1.0)])
        out[f"{name}/q"] = np32(q)
        out[f"{name}/log_alpha"] = np32(ns.marginal_log_mean_coeff(q))
        out[f"{name}/alpha"] = np32(ns.marginal_alpha(q))
        out[f"{name}/sigma"] = np32(ns.marginal_std(q))
        lam = ns.marginal_lambda(q)
        out[f"{name}/lambda"] = np32(lam)
        lq = torch.cat([lam[torch.isfinite(lam)], torch.tensor([-6.5, -5.2, 0.0, 3.3, 5.5, 9.0])])
        out[f"{name}/lq"] = np32(lq)
        out[f"{name}/inv_lambda"] = np32(ns.inverse_lambda(lq))
        solver = DPM_Solver(lambda x, t: x, ns)
        t0 = 1. / ns.total_N
        for skip in ("time_uniform", "logSNR", "time_quadratic"):
            for N in (5, 15, 20, 50):
                out[f"{name}/grid/{skip}/{N}"] = np32(solver.get_time_steps(skip, ns.T, t0, N, "cpu"))
        for steps in (6, 7, 8, 15, 20):
            for order in (1, 2, 3):
                for skip in ("time_uniform", "logSNR"):
                    ts, orders = solver.get_orders_and_timesteps_for_singlestep_solver(steps, order, skip, ns.T, t0, "cpu")
                    out[f"{name}/ss/{skip}/{steps}/{order}/t"] = np32(ts)
                    out[f"{name}/ss/{skip}/{steps}/{order}/orders"] = np.asarray(orders, dtype=np.int64)
        out[f"{name}/total_N"] = np.asarray(ns.total_N)
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), **out)
    return len(out)


def gen_updates():
    out = {}
    x, m0, m1, m2 = (seeded(UPDATE_SHAPE, 100 + i) for i in range(4))
    out["x"], out["m0"], out["m1"], out["m2"] = map(np32, (x, m0, m1, m2))
    lin = lambda xx, tt: 0.3 * xx - 0.1   # model_fn(x, t_continuous) used by the singlestep updates
    for sname in ("sd", "vp_linear"):
        ns = build_schedule(sname)
        ts = torch.linspace(ns.T, 1. / ns.total_N, 21)
        for algo in ("dpmsolver++", "dpmsolver"):
            s = DPM_Solver(lin, ns, algorithm_type=algo)
            for i in (3, 10, 19):   # step index of the target time
                k = f"{sname}/{algo}/{i}"
                out[f"{k}/first"] = np32(s.dpm_solver_first_update(x, ts[i - 1], ts[i], model_s=m0))
                for st in ("dpmsolver", "taylor"):
                    out[f"{k}/ms2/{st}"] = np32(s.multistep_dpm_solver_second_update(
                        x, [m1, m0], [ts[i - 2], ts[i - 1]], ts[i], solver_type=st))
                    out[f"{k}/ms3/{st}"] = np32(s.multistep_dpm_solver_third_update(
                        x, [m2, m1, m0], [ts[i - 3], ts[i - 2], ts[i - 1]], ts[i], solver_type=st))
                    xt, inter = s.singlestep_dpm_solver_second_update(x, ts[i - 1], ts[i], return_intermediate=True, solver_type=st)
                    out[f"{k}/ss2/{st}"] = np32(xt)
                    out[f"{k}/ss2/{st}/model_s1"] = np32(inter["model_s1"])
                    xt, inter = s.singlestep_dpm_solver_third_update(x, ts[i - 1], ts[i], return_intermediate=True, solver_type=st)
                    out[f"{k}/ss3/{st}"] = np32(xt)
                    out[f"{k}/ss3/{st}/model_s2"] = np32(inter["model_s2"])
                    # tensor-valued r1/r2 as sample() passes them (:1223-1227)
                    inner = s.get_time_steps("time_uniform", ts[i - 1].item(), ts[i].item(), 3, "cpu")
                    lam = ns.marginal_lambda(inner)
                    h = lam[-1] - lam[0]
                    r1, r2 = (lam[1] - lam[0]) / h, (lam[2] - lam[0]) / h
                    out[f"{k}/ss3r/{st}"] = np32(s.singlestep_dpm_solver_third_update(
                        x, ts[i - 1], ts[i], r1=r1, r2=r2, solver_type=st))
                    out[f"{k}/ss2r/{st}"] = np32(s.singlestep_dpm_solver_second_update(
                        x, ts[i - 1], ts[i], r1=r1, solver_type=st))
    np.savez_compressed(os.path.join(HERE, "updates.npz"), **out)
    return len(out)


def gen_glue():
    out = {}
    ns = build_schedule("sd")
    B, shape = 4, (4, 16, 16)
    x = seeded((B, *shape), 200)
    bank = seeded((2 * B, *shape), 201)
    t = torch.tensor([0.6004])
    out["x"], out["bank"], out["t"] = np32(x), np32(bank), np32(t)
    # parameterisations (:288-298)
    for mt in ("noise", "x_start", "v", "score"):
        fn = model_wrapper(lambda xx, tt: bank[:B], ns, model_type=mt)
        out[f"param/{mt}"] = np32(fn(x, t.expand(B)))
    # classifier-free guidance (:322-330): network sees the doubled batch, uncond first
    for mt in ("noise", "v"):
        fn = model_wrapper(lambda xx, tt, c: bank, ns, model_type=mt, guidance_type="classifier-free",
                           condition=torch.ones(B, 1), unconditional_condition=torch.zeros(B, 1), guidance_scale=7.5)
        out[f"cfg/{mt}"] = np32(fn(x, t.expand(B)))
    # data prediction (:433-442) without / with dynamic thresholding
    for scale, tag in ((1.0, "big"), (0.05, "small")):   # 'small': quantile < 1 -> max_val floor (:423)
        xs = x * scale
        model = lambda xx, tt: bank[:B] * scale
        s = DPM_Solver(model_wrapper(model, ns), ns, algorithm_type="dpmsolver++")
        out[f"x0/{tag}"] = np32(s.data_prediction_fn(xs, t))
        s = DPM_Solver(model_wrapper(model, ns), ns, algorithm_type="dpmsolver++", correcting_x0_fn="dynamic_thresholding")
        out[f"x0_thr/{tag}"] = np32(s.data_prediction_fn(xs, t))
    # quantile on tiny samples (large gaps between order statistics: exercises the lerp rounding)
    tiny = seeded((1024, 24), 202)
    s = DPM_Solver(lambda xx, tt: xx, ns, correcting_x0_fn="dynamic_thresholding")
    out["tiny"] = np32(tiny)
    out["tiny_thr"] = np32(s.dynamic_thresholding_fn(tiny * 3.0, None))
    out["tiny_q"] = np32(torch.quantile(torch.abs(tiny * 3.0), 0.995, dim=1))
    big = seeded((3, 3 * 64 * 64), 203)
    out["big_q"] = np32(torch.quantile(torch.abs(big), 0.995, dim=1))
    out["big_q_seed"] = np.asarray(203)
    # add_noise (:1012-1030)
    s = DPM_Solver(lambda xx, tt: xx, ns)
    noise = seeded((2, B, *shape), 204)
    out["add_noise"] = np32(s.add_noise(x, torch.tensor([0.3, 0.8]), noise=noise))
    out["add_noise_in"] = np32(noise)
    out["seed_check"] = np32(seeded(16, 1234))   # detects a change of torch's CPU generator
    np.savez_compressed(os.path.join(HERE, "glue.npz"), **out)
    return len(out)


def gen_samples():
    out, meta = {}, []
    for case in SAMPLE_CASES:
        ns = build_schedule(case["schedule"])
        B = case["shape"][0]
        x = seeded(case["shape"], case["seed"])
        calls = []
        net0 = sin_net if case["net"] == "sin" else exact_net

        if case.get("cfg"):
            def net(xx, tt, cc, _n=net0):
                calls.append((float(tt[0]), tuple(xx.shape)))
                return _n(xx, tt) + 0.05 * cc.reshape(-1, 1, 1, 1)
            fn = model_wrapper(net, ns, model_type=case["model_type"], guidance_type="classifier-free",
                               condition=torch.ones(B, 1), unconditional_condition=torch.zeros(B, 1),
                               guidance_scale=case["cfg"])
        else:
            def net(xx, tt, _n=net0):
                calls.append((float(tt[0]), tuple(xx.shape)))
                return _n(xx, tt)
            fn = model_wrapper(net, ns, model_type=case["model_type"])
        s = DPM_Solver(fn, ns, algorithm_type=case["algo"],
                       correcting_x0_fn="dynamic_thresholding" if case.get("thresholding") else None)
        y, inter = s.sample(x, steps=case["steps"], order=case["order"], skip_type=case["skip_type"],
                            method=case["method"], lower_order_final=case.get("lower_order_final", True),
                            denoise_to_zero=case.get("denoise_to_zero", False),
                            solver_type=case.get("solver_type", "dpmsolver"), return_intermediate=True,
                            t_end=case.get("t_end"))
        name = case["name"]
        out[f"{name}/y"] = np32(y)
        if case.get("traj"):   # keep the whole trajectory for a few cases
            out[f"{name}/inter"] = np.stack([np32(v) for v in inter])
        out[f"{name}/calls_t"] = np.asarray([c[0] for c in calls], dtype=np.float32)
        out[f"{name}/calls_b"] = np.asarray([c[1][0] for c in calls], dtype=np.int64)
        meta.append({"name": name, "n_calls": len(calls), "mean": float(y.mean()), "absmean": float(y.abs().mean())})
    np.savez_compressed(os.path.join(HERE, "samples.npz"), **out)
    with open(os.path.join(HERE, "samples_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return len(out)


ADAPTIVE_CASES = [
    dict(name="ad23_eps_vp", schedule="vp_linear", algo="dpmsolver", order=3, t_end=1e-3, solver_type="dpmsolver"),
    dict(name="ad12_eps_vp", schedule="vp_linear", algo="dpmsolver", order=2, t_end=1e-3, solver_type="dpmsolver"),
    dict(name="ad23_pp_sd", schedule="sd", algo="dpmsolver++", order=3, t_end=None, solver_type="taylor"),
    dict(name="ad12_pp_sd", schedule="sd", algo="dpmsolver++", order=2, t_end=None, solver_type="dpmsolver"),
]


def gen_adaptive():
    """dpm_solver_adaptive (:956-1010): final sample and NFE (the reference prints it)."""
    import contextlib
    import io
    out = {}
    for c in ADAPTIVE_CASES:
        ns = build_schedule(c["schedule"])
        x = seeded((2, 3, 8, 8), 77)
        s = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type=c["algo"])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            y = s.sample(x, method="adaptive", order=c["order"], t_end=c["t_end"], solver_type=c["solver_type"])
        out[c["name"] + "/y"] = np32(y)
        out[c["name"] + "/nfe"] = np.asarray(int(buf.getvalue().split()[-1]))
    np.savez_compressed(os.path.join(HERE, "adaptive.npz"), **out)
    return len(out)


if __name__ == "__main__":
    print("adaptive", gen_adaptive())
    print("schedules", gen_schedules())
    print("updates", gen_updates())
    print("glue", gen_glue())
    print("samples", gen_samples())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
