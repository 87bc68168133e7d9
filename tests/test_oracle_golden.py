"""Pin the oracle (oracle/dpm_oracle.py) against golden outputs of the unmodified reference.

numpy namespace: independent arithmetic -- exp/log/expm1 may differ from torch by an ulp, so the
tolerance is a few fp32 ulps on scalars and 2e-6 relative on tensors; where no transcendental is
involved (interpolation, linspace, quantile, element-wise forms) it must be bit-exact.
torch-CPU namespace (the one bench.py times as the CPU baseline): bit-exact everywhere."""
import numpy as np
import pytest
import torch

from cases import SAMPLE_CASES, SCHEDULES, make_betas
from helpers import CASES, oracle_schedule, rel_err, run_oracle_case
from oracle import dpm_oracle as O

TH = O.torch_namespace()


@pytest.mark.parametrize("name", SCHEDULES)
def test_schedule_scalars(golden, name):
    g = golden["schedules"]
    ns = oracle_schedule(name)
    if ns.schedule == "discrete":
        np.testing.assert_array_equal(ns.t, g[f"{name}/t_array"])               # linspace emulation is exact
        np.testing.assert_allclose(ns.log_alpha, g[f"{name}/log_alpha_array"], rtol=3e-7, atol=1e-9)
        assert ns.total_N == int(g[f"{name}/total_N"])
        ns.set_tables(g[f"{name}/t_array"], g[f"{name}/log_alpha_array"])
        np.testing.assert_array_equal(ns.marginal_log_mean_coeff(g[f"{name}/q"]), g[f"{name}/log_alpha"])  # no transcendental
    q = g[f"{name}/q"]
    np.testing.assert_allclose(ns.marginal_log_mean_coeff(q), g[f"{name}/log_alpha"], rtol=1e-6)
    np.testing.assert_allclose(ns.marginal_alpha(q), g[f"{name}/alpha"], rtol=3e-7)
    # operating range of the solver is [1/N, T]; below it 1 - exp(2 log alpha) keeps only a few bits
    fin = np.isfinite(g[f"{name}/lambda"]) & (q >= 9e-4)
    # sigma = sqrt(1 - exp(2 log alpha)) cancels catastrophically as t -> 0 (log alpha ~ -1e-5): one
    # ulp of numpy's exp() vs torch's moves sigma by ~1e-5 absolute and lambda = log alpha - log sigma
    # by ~2e-3 there. (The torch-namespace test below is bit-exact on the same points.)
    np.testing.assert_allclose(ns.marginal_std(q)[fin], g[f"{name}/sigma"][fin], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(ns.marginal_lambda(q)[fin], g[f"{name}/lambda"][fin], rtol=2e-4, atol=3e-3)
    np.testing.assert_allclose(ns.inverse_lambda(g[f"{name}/lq"]), g[f"{name}/inv_lambda"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", SCHEDULES)
def test_schedule_scalars_torch_namespace_bit_exact(golden, name):
    g = golden["schedules"]
    ns = oracle_schedule(name, xp=TH)
    q = torch.from_numpy(g[f"{name}/q"])
    if ns.schedule == "discrete":
        assert torch.equal(ns.log_alpha, torch.from_numpy(g[f"{name}/log_alpha_array"]))
    for fn, key in ((ns.marginal_log_mean_coeff, "log_alpha"), (ns.marginal_alpha, "alpha"), (ns.marginal_std, "sigma"),
                    (ns.marginal_lambda, "lambda")):
        np.testing.assert_array_equal(fn(q).numpy(), g[f"{name}/{key}"])
    np.testing.assert_array_equal(ns.inverse_lambda(torch.from_numpy(g[f"{name}/lq"])).numpy(), g[f"{name}/inv_lambda"])


@pytest.mark.parametrize("name", SCHEDULES)
def test_time_grids_and_orders(golden, name):
    g = golden["schedules"]
    ns = oracle_schedule(name, g)
    t0 = 1. / ns.total_N
    for N in (5, 15, 20, 50):
        np.testing.assert_array_equal(O.time_steps(ns, "time_uniform", ns.T, t0, N), g[f"{name}/grid/time_uniform/{N}"])
        np.testing.assert_array_equal(O.time_steps(ns, "time_quadratic", ns.T, t0, N), g[f"{name}/grid/time_quadratic/{N}"])
        np.testing.assert_allclose(O.time_steps(ns, "logSNR", ns.T, t0, N), g[f"{name}/grid/logSNR/{N}"], rtol=3e-4, atol=1e-6)
    for steps in (6, 7, 8, 15, 20):
        for order in (1, 2, 3):
            assert O.singlestep_orders(steps, order) == g[f"{name}/ss/time_uniform/{steps}/{order}/orders"].tolist()


def _update_cases(g, ns, algo, xp, conv):
    x, m0, m1, m2 = (conv(g[k]) for k in ("x", "m0", "m1", "m2"))
    ts = xp.linspace(ns.T, 1. / ns.total_N, 21)
    net = lambda xx, tt: 0.3 * xx - 0.1
    # DPM_Solver.model_fn: the network predicts noise; dpmsolver++ buffers x0 (:444-451)
    lin = (lambda xx, tt: O.data_prediction(ns, xx, net(xx, tt), tt)) if algo == "dpmsolver++" else net
    for i in (3, 10, 19):
        T = lambda j: ts[j:j + 1]
        yield f"{i}/first", O.first_update(ns, algo, x, T(i - 1), T(i), m0)
        for st in ("dpmsolver", "taylor"):
            yield f"{i}/ms2/{st}", O.multistep_second(ns, algo, st, x, [m1, m0], [T(i - 2), T(i - 1)], T(i))
            yield f"{i}/ms3/{st}", O.multistep_third(ns, algo, x, [m2, m1, m0], [T(i - 3), T(i - 2), T(i - 1)], T(i))
            yield f"{i}/ss2/{st}", O.singlestep_second(ns, algo, st, x, T(i - 1), T(i), lin)[0]
            yield f"{i}/ss3/{st}", O.singlestep_third(ns, algo, st, x, T(i - 1), T(i), lin)[0]


@pytest.mark.parametrize("sname", ["sd", "vp_linear"])
@pytest.mark.parametrize("algo", ["dpmsolver++", "dpmsolver"])
def test_updates(golden, sname, algo):
    g = golden["updates"]
    ns = oracle_schedule(sname, golden["schedules"])
    for key, got in _update_cases(g, ns, algo, O.NP, lambda a: a):
        assert rel_err(got, g[f"{sname}/{algo}/{key}"]) < 2e-5, key   # numpy expm1 ulp x phi_3 cancellation (SURVEY hard part 1)
    ns = oracle_schedule(sname, xp=TH)
    for key, got in _update_cases(g, ns, algo, TH, torch.from_numpy):
        np.testing.assert_array_equal(got.numpy(), g[f"{sname}/{algo}/{key}"], err_msg=key)


def test_glue(golden):
    g = golden["glue"]
    ns = oracle_schedule("sd", golden["schedules"])
    x, bank, t = g["x"], g["bank"], g["t"]
    B = x.shape[0]
    for mt in ("noise", "x_start", "v", "score"):
        assert rel_err(O.to_noise(ns, mt, x, bank[:B], t), g[f"param/{mt}"]) < 1e-6
    np.testing.assert_array_equal(O.cfg_combine(bank[:B], bank[B:], np.float32(7.5)), g["cfg/noise"])
    both = O.to_noise(ns, "v", np.concatenate([x, x]), bank, t)
    assert rel_err(O.cfg_combine(both[:B], both[B:], np.float32(7.5)), g["cfg/v"]) < 1e-6
    for scale, tag in ((np.float32(1.0), "big"), (np.float32(0.05), "small")):
        assert rel_err(O.data_prediction(ns, x * scale, bank[:B] * scale, t), g[f"x0/{tag}"]) < 1e-6
        assert rel_err(O.data_prediction(ns, x * scale, bank[:B] * scale, t, (0.995, 1.0)), g[f"x0_thr/{tag}"]) < 1e-6
    # quantile: bit-exact, including the fused lerp rounding
    np.testing.assert_array_equal(O.quantile_abs(g["tiny"] * np.float32(3.0), 0.995), g["tiny_q"])
    np.testing.assert_array_equal(O.dynamic_thresholding(g["tiny"] * np.float32(3.0)), g["tiny_thr"])
    from cases import seeded
    np.testing.assert_array_equal(O.quantile_abs(seeded((3, 3 * 64 * 64), 203).numpy(), 0.995), g["big_q"])
    np.testing.assert_array_equal(seeded(16, 1234).numpy(), g["seed_check"])


@pytest.mark.parametrize("name", [c["name"] for c in SAMPLE_CASES])
def test_sample_loops(golden, name):
    """The oracle's own loops (multistep / singlestep) reproduce the reference: outputs within
    1e-4 relative (numpy transcendental ulps propagate through <= 20 steps; the torch-namespace test
    below is bit-exact), identical call trace."""
    g = golden["samples"]
    case = CASES[name]
    y, inter, calls = run_oracle_case(case, golden["schedules"])
    assert [c[1][0] for c in calls] == g[f"{name}/calls_b"].tolist()
    np.testing.assert_allclose(np.asarray([c[0] for c in calls]), g[f"{name}/calls_t"], rtol=2e-4, atol=1e-3)
    assert rel_err(y, g[f"{name}/y"]) < 1e-4


@pytest.mark.parametrize("name", ["pp2m", "pp3m", "eps3s", "eps3s_cfg", "pp2m_logsnr", "pp3s_taylor", "eps2m_taylor", "pp2m_v"])
def test_sample_loops_torch_namespace_bit_exact(golden, name):
    g = golden["samples"]
    y, _, calls = run_oracle_case(CASES[name], xp=TH)
    np.testing.assert_array_equal(y.numpy(), g[f"{name}/y"])
    np.testing.assert_array_equal(np.asarray([c[0] for c in calls], dtype=np.float32), g[f"{name}/calls_t"])
