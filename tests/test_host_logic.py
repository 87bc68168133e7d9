"""CPU tests of the product's host side: schedule scalars, time grids, the coefficient plan and the
whole sample() control flow, executed on the numpy executor (tests/oracle_backend.py) and compared
BIT FOR BIT with golden outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from cases import SAMPLE_CASES, SCHEDULES, make_betas
from helpers import CASES, product_schedule, run_product_case


def t(a):
    return torch.from_numpy(np.asarray(a))


# ---- T0: schedule scalars -------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCHEDULES)
def test_schedule_scalars_bit_exact(golden, name):
    g = golden["schedules"]
    ns = product_schedule(name)
    q = t(g[f"{name}/q"])
    if ns.schedule == "discrete":
        assert ns.total_N == int(g[f"{name}/total_N"])
        assert torch.equal(ns.t_array.reshape(-1), t(g[f"{name}/t_array"]))
        assert torch.equal(ns.log_alpha_array.reshape(-1), t(g[f"{name}/log_alpha_array"]))
    np.testing.assert_array_equal(ns.marginal_log_mean_coeff(q).numpy(), g[f"{name}/log_alpha"])
    np.testing.assert_array_equal(ns.marginal_alpha(q).numpy(), g[f"{name}/alpha"])
    np.testing.assert_array_equal(ns.marginal_std(q).numpy(), g[f"{name}/sigma"])
    np.testing.assert_array_equal(ns.marginal_lambda(q).numpy(), g[f"{name}/lambda"])
    np.testing.assert_array_equal(ns.inverse_lambda(t(g[f"{name}/lq"])).numpy(), g[f"{name}/inv_lambda"])


def test_cosine_clip_and_known_values():
    ns = product_schedule("iddpm_cosine")
    assert ns.total_N == 996
    lam = float(ns.marginal_lambda(torch.tensor([1.0])))
    assert abs(lam - (-5.0778)) < 1e-3
    ns = product_schedule("ddpm_linear")
    assert ns.total_N == 1000
    assert abs(float(ns.marginal_lambda(torch.tensor([1.0]))) + 5.0588) < 1e-3
    assert abs(float(ns.marginal_lambda(torch.tensor([1e-3]))) - 4.6050) < 1e-3


def test_interpolate_fn_matches_reference_probe():
    from dpm_solver_b200 import interpolate_fn
    y = interpolate_fn(torch.tensor([[-0.5], [0.25], [2.0]]), torch.tensor([[0., 1., 1.5]]), torch.tensor([[0., 10., 11.]]))
    assert y.reshape(-1).tolist() == [-5.0, 2.5, 12.0]        # SURVEY 8a probe


def test_schedule_errors():
    from dpm_solver_b200 import NoiseScheduleVP
    with pytest.raises(ValueError):
        NoiseScheduleVP("quadratic")
    with pytest.raises(AssertionError):
        NoiseScheduleVP("discrete")


@pytest.mark.parametrize("name", SCHEDULES)
def test_time_grids_bit_exact(golden, name):
    from dpm_solver_b200 import DPM_Solver
    g = golden["schedules"]
    ns = product_schedule(name)
    s = DPM_Solver(lambda x, tt: x, ns)
    t0 = 1. / ns.total_N
    for skip in ("time_uniform", "logSNR", "time_quadratic"):
        for N in (5, 15, 20, 50):
            np.testing.assert_array_equal(s.get_time_steps(skip, ns.T, t0, N, "cpu").numpy(), g[f"{name}/grid/{skip}/{N}"])
    for steps in (6, 7, 8, 15, 20):
        for order in (1, 2, 3):
            for skip in ("time_uniform", "logSNR"):
                ts, orders = s.get_orders_and_timesteps_for_singlestep_solver(steps, order, skip, ns.T, t0, "cpu")
                np.testing.assert_array_equal(ts.numpy(), g[f"{name}/ss/{skip}/{steps}/{order}/t"])
                assert list(orders) == g[f"{name}/ss/{skip}/{steps}/{order}/orders"].tolist()
    assert s.get_orders_and_timesteps_for_singlestep_solver(15, 3, "time_uniform", 1., 1e-3, "cpu")[1] == [3, 3, 3, 3, 2, 1]
    with pytest.raises(ValueError):
        s.get_time_steps("cubic", 1., 1e-3, 5, "cpu")
    with pytest.raises(ValueError):
        s.get_orders_and_timesteps_for_singlestep_solver(10, 4, "time_uniform", 1., 1e-3, "cpu")


def test_known_coefficients():
    """2M++/3M++ scalars of step 10 -> 11 on the SD schedule."""
    from dpm_solver_b200 import plan as P
    ns = product_schedule("sd")
    ts = torch.linspace(1., 1e-3, 21)
    pl2 = P.multistep_plan(ns, "dpmsolver++", "dpmsolver", ts, 2, True)
    c = pl2[10]
    assert abs(c.a - 0.94983894) < 2e-7 and abs(-c.c0 - (-0.08976490)) < 2e-7 and abs(1 / c.w0 - 1.02388406) < 2e-6
    pl3 = P.multistep_plan(ns, "dpmsolver++", "dpmsolver", ts, 3, True)
    c = pl3[10]
    alpha_t = float(ns.marginal_alpha(ts[11:12]))
    assert abs(-c.c0 / alpha_t - (-0.15222831)) < 1e-6
    assert abs(c.c1 / alpha_t - 0.07820815) < 1e-6
    assert abs(-c.c2 / alpha_t - (-0.02642426)) < 1e-6
    assert [p.order for p in pl3[:4]] == [1, 2, 3, 3]
    assert P.multistep_orders(8, 3, True) == [1, 2, 3, 3, 3, 3, 2, 1]
    assert P.multistep_orders(8, 3, False) == [1, 2, 3, 3, 3, 3, 3, 3]


# ---- T1: single updates through the public methods -------------------------------------------------
@pytest.mark.parametrize("sname", ["sd", "vp_linear"])
@pytest.mark.parametrize("algo", ["dpmsolver++", "dpmsolver"])
def test_update_methods_bit_exact(golden, oracle_backend, sname, algo):
    from dpm_solver_b200 import DPM_Solver
    g = golden["updates"]
    x, m0, m1, m2 = (t(g[k]) for k in ("x", "m0", "m1", "m2"))
    ns = product_schedule(sname)
    s = DPM_Solver(lambda xx, tt: 0.3 * xx - 0.1, ns, algorithm_type=algo)
    ts = torch.linspace(ns.T, 1. / ns.total_N, 21)
    for i in (3, 10, 19):
        k = f"{sname}/{algo}/{i}"
        np.testing.assert_array_equal(s.dpm_solver_first_update(x, ts[i - 1], ts[i], model_s=m0).numpy(), g[f"{k}/first"])
        np.testing.assert_array_equal(
            s.multistep_dpm_solver_update(x, [m1, m0], [ts[i - 2], ts[i - 1]], ts[i], 1).numpy(), g[f"{k}/first"])
        for st in ("dpmsolver", "taylor"):
            got = s.multistep_dpm_solver_second_update(x, [m1, m0], [ts[i - 2], ts[i - 1]], ts[i], solver_type=st)
            np.testing.assert_array_equal(got.numpy(), g[f"{k}/ms2/{st}"])
            got = s.multistep_dpm_solver_third_update(x, [m2, m1, m0], [ts[i - 3], ts[i - 2], ts[i - 1]], ts[i], solver_type=st)
            np.testing.assert_array_equal(got.numpy(), g[f"{k}/ms3/{st}"])
            xt, inter = s.singlestep_dpm_solver_second_update(x, ts[i - 1], ts[i], return_intermediate=True, solver_type=st)
            np.testing.assert_array_equal(xt.numpy(), g[f"{k}/ss2/{st}"])
            np.testing.assert_array_equal(inter["model_s1"].numpy(), g[f"{k}/ss2/{st}/model_s1"])
            xt, inter = s.singlestep_dpm_solver_third_update(x, ts[i - 1], ts[i], return_intermediate=True, solver_type=st)
            np.testing.assert_array_equal(xt.numpy(), g[f"{k}/ss3/{st}"])
            np.testing.assert_array_equal(inter["model_s2"].numpy(), g[f"{k}/ss3/{st}/model_s2"])
            inner = s.get_time_steps("time_uniform", ts[i - 1].item(), ts[i].item(), 3, "cpu")
            lam = ns.marginal_lambda(inner)
            h = lam[-1] - lam[0]
            r1, r2 = (lam[1] - lam[0]) / h, (lam[2] - lam[0]) / h
            got = s.singlestep_dpm_solver_update(x, ts[i - 1], ts[i], 3, solver_type=st, r1=r1, r2=r2)
            np.testing.assert_array_equal(got.numpy(), g[f"{k}/ss3r/{st}"])
            got = s.singlestep_dpm_solver_update(x, ts[i - 1], ts[i], 2, solver_type=st, r1=r1)
            np.testing.assert_array_equal(got.numpy(), g[f"{k}/ss2r/{st}"])
    with pytest.raises(ValueError):
        s.multistep_dpm_solver_second_update(x, [m1, m0], [ts[1], ts[2]], ts[3], solver_type="euler")
    with pytest.raises(ValueError):
        s.singlestep_dpm_solver_update(x, ts[1], ts[2], 4)
    with pytest.raises(ValueError):
        s.multistep_dpm_solver_update(x, [m1, m0], [ts[1], ts[2]], ts[3], 5)


# ---- T2: glue ------------------------------------------------------------------------------------
def test_glue_bit_exact(golden, oracle_backend):
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    g = golden["glue"]
    ns = product_schedule("sd")
    x, bank, tt = t(g["x"]), t(g["bank"]), t(g["t"])
    B = x.shape[0]
    for mt in ("noise", "x_start", "v", "score"):
        fn = model_wrapper(lambda xx, t_: bank[:B], ns, model_type=mt)
        np.testing.assert_array_equal(fn(x, tt.expand(B)).numpy(), g[f"param/{mt}"])
    for mt in ("noise", "v"):
        fn = model_wrapper(lambda xx, t_, c: bank, ns, model_type=mt, guidance_type="classifier-free",
                           condition=torch.ones(B, 1), unconditional_condition=torch.zeros(B, 1), guidance_scale=7.5)
        np.testing.assert_array_equal(fn(x, tt.expand(B)).numpy(), g[f"cfg/{mt}"])
    for scale, tag in ((1.0, "big"), (0.05, "small")):
        xs = x * scale
        model = lambda xx, t_: bank[:B] * scale
        s = DPM_Solver(model_wrapper(model, ns), ns)
        np.testing.assert_array_equal(s.data_prediction_fn(xs, tt).numpy(), g[f"x0/{tag}"])
        s = DPM_Solver(model_wrapper(model, ns), ns, correcting_x0_fn="dynamic_thresholding")
        np.testing.assert_array_equal(s.data_prediction_fn(xs, tt).numpy(), g[f"x0_thr/{tag}"])
    s = DPM_Solver(lambda xx, t_: xx, ns, correcting_x0_fn="dynamic_thresholding")
    np.testing.assert_array_equal(s.dynamic_thresholding_fn(t(g["tiny"]) * 3.0, None).numpy(), g["tiny_thr"])
    s = DPM_Solver(lambda xx, t_: xx, ns)
    got = s.add_noise(x, torch.tensor([0.3, 0.8]), noise=t(g["add_noise_in"]))
    np.testing.assert_array_equal(got.numpy(), g["add_noise"])


def test_wrapper_asserts():
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule("sd")
    with pytest.raises(AssertionError):
        model_wrapper(lambda x, t: x, ns, model_type="epsilon")
    with pytest.raises(AssertionError):
        model_wrapper(lambda x, t: x, ns, guidance_type="cfg")
    with pytest.raises(AssertionError):
        DPM_Solver(lambda x, t: x, ns, algorithm_type="ddim")


# ---- T3: whole loops ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [c["name"] for c in SAMPLE_CASES])
def test_sample_bit_exact_and_call_order(golden, oracle_backend, name):
    """sample() on the numpy executor == the reference, bit for bit, with the identical sequence of
    network calls (time label and batch shape of every call)."""
    g = golden["samples"]
    case = CASES[name]
    y, inter, calls = run_product_case(case, device="cpu")
    np.testing.assert_array_equal(np.asarray([c[0] for c in calls], dtype=np.float32), g[f"{name}/calls_t"])
    assert [c[1][0] for c in calls] == g[f"{name}/calls_b"].tolist()
    np.testing.assert_array_equal(y.numpy(), g[f"{name}/y"])
    if f"{name}/inter" in g:
        ref = g[f"{name}/inter"]
        assert len(inter) == ref.shape[0]
        for a, b in zip(inter, ref):
            np.testing.assert_array_equal(a.numpy(), b)
        assert len({v.data_ptr() for v in inter}) == len(inter)      # distinct tensors, never in place


def test_launch_budget(oracle_backend):
    """One fused launch per model evaluation (plus one quantile launch with thresholding; plus, under CFG, the one
    cat([x] * 2) of a run's first evaluation -- later ones are written by the update kernel itself)."""
    _, _, calls = run_product_case(CASES["pp2m"], device="cpu")
    assert oracle_backend.launches == 20 == len(calls)
    oracle_backend.launches = 0
    _, _, calls = run_product_case(CASES["eps3s_cfg"], device="cpu")
    assert oracle_backend.launches == 15 + 1 and len(calls) == 15
    oracle_backend.launches = 0
    run_product_case(CASES["pp3m_thr"], device="cpu")
    assert oracle_backend.launches == 40


def test_sample_asserts_and_errors(oracle_backend):
    from dpm_solver_b200 import DPM_Solver
    ns = product_schedule("sd")
    s = DPM_Solver(lambda x, tt: 0.1 * x, ns)
    x = torch.randn(2, 4, 8, 8)
    with pytest.raises(AssertionError):
        s.sample(x, steps=2, order=3)                      # steps >= order (:1172)
    with pytest.raises(ValueError):
        s.sample(x, steps=5, method="rk45")                # :1234
    with pytest.raises(AssertionError):
        s.sample(x, steps=5, t_end=0.0)                    # :1161
    with pytest.raises(AssertionError):
        s.sample(x, method="adaptive", return_intermediate=True)   # :1163
    with pytest.raises(ValueError):
        s.sample(x, steps=5, skip_type="cubic")
    with pytest.raises(TypeError):
        s.sample(x.double(), steps=5)


def test_hooks_and_custom_corrector(oracle_backend):
    """correcting_xt_fn sees (x, t, step) after every update; a user correcting_x0_fn sees x0."""
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule("sd")
    seen = []

    def fix_xt(x, tt, step):
        seen.append((float(tt), step))
        return x * 0.99

    clip = lambda x0, tt: x0.clamp(-1, 1)
    net = lambda x, tt: 0.1 * x
    s = DPM_Solver(model_wrapper(net, ns), ns, correcting_x0_fn=clip, correcting_xt_fn=fix_xt)
    x = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    y = s.sample(x, steps=6, order=2)
    assert [st for _, st in seen] == list(range(0, 7))
    assert torch.isfinite(y).all()


def test_plan_cache_follows_the_schedule(oracle_backend):
    """The cached coefficient plan is keyed on the schedule tables: replacing or editing them in
    place must not serve a stale plan; repeated calls with one configuration reuse it."""
    from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP, model_wrapper
    from cases import exact_net, make_betas, seeded
    ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("sd")[1]))
    s = DPM_Solver(model_wrapper(exact_net, ns), ns)
    x = seeded((2, 4, 8, 8), 9)
    y1 = s.sample(x, steps=10, order=2)
    assert len(s._plan_cache) == 1
    y1b = s.sample(x, steps=10, order=2)
    assert len(s._plan_cache) == 1 and torch.equal(y1, y1b)
    s.sample(x, steps=10, order=3)
    assert len(s._plan_cache) == 2
    # other schedule tables -> a fresh solver on them gives the same answer as the cached one must now give
    ns2 = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("ddpm_linear")[1]))
    ns.log_alpha_array, ns.t_array, ns.total_N = ns2.log_alpha_array, ns2.t_array, ns2.total_N
    y2 = s.sample(x, steps=10, order=2)
    ref = DPM_Solver(model_wrapper(exact_net, ns2), ns2).sample(x, steps=10, order=2)
    assert torch.equal(y2, ref) and not torch.equal(y2, y1)
    ns.log_alpha_array.mul_(1.01)                     # in-place edit
    y3 = s.sample(x, steps=10, order=2)
    assert not torch.equal(y3, y2)


def test_older_constructor_keywords(oracle_backend):
    """predict_x0 / thresholding / max_val (the JAX twin's constructor, dpm_solver_jax.py:351) select the
    same solver as algorithm_type / correcting_x0_fn / thresholding_max_val."""
    from cases import exact_net, seeded
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule("ddpm_linear")
    x = seeded((2, 3, 8, 8), 3)
    fn = model_wrapper(exact_net, ns)
    new = DPM_Solver(fn, ns, algorithm_type="dpmsolver++", correcting_x0_fn="dynamic_thresholding", thresholding_max_val=1.5)
    old = DPM_Solver(fn, ns, predict_x0=True, thresholding=True, max_val=1.5)
    assert old.algorithm_type == "dpmsolver++" and old.thresholding_max_val == 1.5
    np.testing.assert_array_equal(old.sample(x, steps=6, order=2).numpy(), new.sample(x, steps=6, order=2).numpy())
    assert DPM_Solver(fn, ns, predict_x0=False).algorithm_type == "dpmsolver"


def test_non_fp32_schedule_is_rejected():
    """NoiseScheduleVP(dtype=float64) makes the reference promote x and every update to fp64; there are no
    fp64 kernels, so the solver refuses instead of silently answering in fp32."""
    from dpm_solver_b200 import DPM_Solver, NoiseScheduleVP
    ns = NoiseScheduleVP("discrete", betas=torch.from_numpy(make_betas("sd")[1]), dtype=torch.float64)
    assert ns.log_alpha_array.dtype == torch.float64        # the schedule object itself follows the reference
    with pytest.raises(TypeError, match="fp32"):
        DPM_Solver(lambda x, t: x, ns)


def test_one_argument_x0_corrector(oracle_backend):
    """The older vendored solver copy calls `correcting_x0_fn(x0)` (examples/stable-diffusion/.../dpm_solver.py
    :447-448): a one-argument callable is accepted and gives the same result as its two-argument spelling."""
    from cases import exact_net, seeded
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    ns = product_schedule("sd")
    x = seeded((2, 3, 8, 8), 9)
    fn = model_wrapper(exact_net, ns)
    y1 = DPM_Solver(fn, ns, correcting_x0_fn=lambda x0: x0.clamp(-1, 1)).sample(x, steps=5, order=2)
    y2 = DPM_Solver(fn, ns, correcting_x0_fn=lambda x0, t: x0.clamp(-1, 1)).sample(x, steps=5, order=2)
    np.testing.assert_array_equal(y1.numpy(), y2.numpy())
