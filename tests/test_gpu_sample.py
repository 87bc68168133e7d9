"""GPU loop parity: DPM_Solver.sample() on cuda:0 through the C-ABI against golden outputs of the
unmodified reference (fp32: bit-identical for networks made of exact IEEE ops; <=1e-5 relative, the
north-star tolerance, when the network itself calls sin() on the device)."""
import numpy as np
import pytest
import torch

from cases import SAMPLE_CASES
from helpers import CASES, rel_err, run_product_case

pytestmark = pytest.mark.gpu
TOL = 1e-5   # BASELINE.json north_star: "<=1e-5 relative fp32"


@pytest.mark.parametrize("name", [c["name"] for c in SAMPLE_CASES])
def test_sample_matches_reference(golden, cuda_backend, name):
    g = golden["samples"]
    case = CASES[name]
    before = cuda_backend.launch_count()
    y, inter, calls = run_product_case(case, device="cuda:0")
    assert cuda_backend.launch_count() > before, "the CUDA library did not run"
    assert y.is_cuda and y.dtype == torch.float32
    # identical sequence of network calls
    np.testing.assert_allclose(np.asarray([c[0] for c in calls], dtype=np.float32), g[f"{name}/calls_t"], rtol=1e-6)
    assert [c[1][0] for c in calls] == g[f"{name}/calls_b"].tolist()
    if case["net"] == "exact":
        np.testing.assert_array_equal(y.cpu().numpy(), g[f"{name}/y"])
        if f"{name}/inter" in g:
            for a, b in zip(inter, g[f"{name}/inter"]):
                np.testing.assert_array_equal(a.cpu().numpy(), b)
    else:
        assert rel_err(y.cpu().numpy(), g[f"{name}/y"]) <= TOL


def test_known_answers_config1(golden, cuda_backend):
    """DPM-Solver++2M, 20 steps, [8,4,64,64], sin network."""
    y, _, calls = run_product_case(CASES["c1_pp2m_sin"], device="cuda:0")
    assert len(calls) == 20 and abs(calls[0][0] - 999.0) < 1e-3 and abs(calls[-1][0] - 49.95) < 1e-2
    assert abs(float(y.mean()) - (-0.02713880)) < 1e-5
    assert abs(float(y.abs().mean()) - 8.39973736) < 1e-4
    ref = [-1.21692860, -5.25487947, 1.66610932]
    assert np.allclose(y[0, 0, 0, :3].cpu().numpy(), ref, rtol=1e-5)


# max|y - y_ref| / max|y_ref| of 16-bit storage (state AND network output) against the fp32 reference golden,
# measured with the numpy executor (which the CUDA path reproduces bit for bit): (bf16, f16)
MEASURED_16 = {"pp2m": (1.598e-2, 2.764e-3), "pp3m": (5.049e-2, 5.204e-3), "eps3s_cfg": (5.927e-3, 1.425e-3),
               "pp3m_thr": (5.929e-2, 6.187e-3), "pp2m_cfg_v": (2.751e-2, 2.979e-3), "pp3s_taylor": (1.541e-2, 1.825e-3)}


@pytest.mark.parametrize("name", ["pp2m", "pp3m", "eps3s_cfg", "pp3m_thr", "pp2m_cfg_v", "pp3s_taylor"])
@pytest.mark.parametrize("sdt", [torch.bfloat16, torch.float16])
def test_sample_16bit_state(golden, cuda_backend, oracle_backend_cpu, name, sdt):
    """16-bit storage (fp32 math, RN on store): bitwise equal to the numpy executor run with the
    same storage type, and within 1.5x the MEASURED storage-precision deviation from the fp32 reference."""
    from dpm_solver_b200 import ops
    case = CASES[name]
    y, _, _ = run_product_case(case, device="cuda:0", state_dtype=sdt, model_dtype=sdt)
    assert y.dtype == sdt
    ops.set_backend(oracle_backend_cpu)
    y_ref, _, _ = run_product_case(case, device="cpu", state_dtype=sdt, model_dtype=sdt)
    if not case.get("thresholding"):
        assert torch.equal(y.cpu(), y_ref)
    g = golden["samples"][f"{name}/y"]
    measured = MEASURED_16[name][0 if sdt == torch.bfloat16 else 1]
    assert rel_err(y.float().cpu().numpy(), g) <= 1.5 * measured      # (the full-size bound: test_vs_reference_workloads.py)


def test_mixed_bf16_model_fp32_state(golden, cuda_backend):
    """Reference promotion: a bf16 network output feeds an fp32 state (SURVEY 8a, bf16 note)."""
    case = CASES["pp2m_cfg"]
    y, _, _ = run_product_case(case, device="cuda:0", model_dtype=torch.bfloat16)
    assert y.dtype == torch.float32
    assert rel_err(y.cpu().numpy(), golden["samples"]["pp2m_cfg/y"]) < 0.05


def test_inverse_round_trip(cuda_backend):
    """sample() then inverse() returns to the start for a smooth network (independent property)."""
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    from helpers import product_schedule
    ns = product_schedule("sd")
    net = lambda x, t: 0.3 * x
    s = DPM_Solver(model_wrapper(net, ns), ns, algorithm_type="dpmsolver++")
    x = torch.randn(4, 4, 16, 16, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(0))
    y = s.sample(x, steps=60, order=3, t_end=0.05)
    xr = s.inverse(y, steps=60, order=3, t_start=0.05, t_end=1.0)
    assert rel_err(xr.cpu().numpy(), x.cpu().numpy()) < 2e-3


def test_order1_is_ddim_closed_form(cuda_backend):
    """Order 1 == DDIM: x_t = alpha_t*x0 + sigma_t*eps with x0 from (x, eps) (analytic check)."""
    from dpm_solver_b200 import DPM_Solver
    from helpers import product_schedule
    ns = product_schedule("sd")
    eps = torch.randn(2, 4, 8, 8, device="cuda:0")
    s = DPM_Solver(lambda x, t: eps, ns, algorithm_type="dpmsolver++")
    x = torch.randn(2, 4, 8, 8, device="cuda:0")
    ts, tt = torch.tensor([0.8]), torch.tensor([0.6])
    got = s.dpm_solver_first_update(x, ts.cuda(), tt.cuda())
    a_s, s_s, a_t, s_t = (float(v) for v in (ns.marginal_alpha(ts), ns.marginal_std(ts), ns.marginal_alpha(tt), ns.marginal_std(tt)))
    x0 = (x - s_s * eps) / a_s
    assert rel_err(got.cpu().numpy(), (a_t * x0 + s_t * eps).cpu().numpy()) < 1e-5


def test_adaptive_runs(cuda_backend, capsys):
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    from helpers import product_schedule
    ns = product_schedule("vp_linear")
    s = DPM_Solver(model_wrapper(lambda x, t: 0.2 * x, ns), ns, algorithm_type="dpmsolver")
    x = torch.randn(2, 3, 8, 8, device="cuda:0")
    y = s.sample(x, method="adaptive", order=3, t_end=1e-3)
    assert torch.isfinite(y).all() and "adaptive solver nfe" in capsys.readouterr().out


@pytest.mark.parametrize("name", ["pp2m", "eps3s_cfg", "pp3m_thr"])
def test_whole_sample_loop_is_cuda_graph_capturable(golden, cuda_backend, name):
    """once the plan and the device tables are cached (first call), sample() issues no
    host<->device copy, no sync and no collective, so the whole loop -- network calls included, if
    the network is capturable -- can be captured in one CUDA graph and replayed."""
    from cases import exact_net, make_betas, seeded
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    from helpers import product_schedule
    case = CASES[name]
    ns = product_schedule(case["schedule"])
    B = case["shape"][0]
    x = seeded(case["shape"], case["seed"]).cuda()
    if case.get("cfg"):
        net = lambda xx, tt, cc: exact_net(xx, tt) + 0.05 * cc.reshape(-1, 1, 1, 1)
        fn = model_wrapper(net, ns, guidance_type="classifier-free", condition=torch.ones(B, 1, device="cuda"),
                           unconditional_condition=torch.zeros(B, 1, device="cuda"), guidance_scale=case["cfg"])
    else:
        fn = model_wrapper(exact_net, ns)
    s = DPM_Solver(fn, ns, algorithm_type=case["algo"],
                   correcting_x0_fn="dynamic_thresholding" if case.get("thresholding") else None)
    kw = dict(steps=case["steps"], order=case["order"], skip_type=case["skip_type"], method=case["method"])
    y_eager = s.sample(x, **kw)                      # warms the plan / table caches
    x_static = x.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        s.sample(x_static, **kw)                     # warm-up on the capture stream
        with torch.cuda.graph(g, stream=side):
            y_static = s.sample(x_static, **kw)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_static, y_eager)
    np.testing.assert_array_equal(y_static.cpu().numpy(), golden["samples"][f"{name}/y"])
    x_static.copy_(x * 0.5)                          # new input, same graph
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_static, s.sample(x * 0.5, **kw))


def test_capture_api(golden, cuda_backend):
    """DPM_Solver.capture(): one graph launch per sample() -- zero per-step host work."""
    from cases import exact_net, seeded
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    from helpers import product_schedule
    ns = product_schedule("sd")
    s = DPM_Solver(model_wrapper(exact_net, ns), ns)
    x = seeded((2, 4, 16, 16), 1234).cuda()
    run = s.capture(x, steps=20, order=2)
    before = cuda_backend.launch_count()
    y = run(x).clone()
    assert cuda_backend.launch_count() == before          # replay launches nothing through the host path
    np.testing.assert_array_equal(y.cpu().numpy(), golden["samples"]["pp2m/y"])
    y2 = run(x * 0.25).clone()
    assert torch.equal(y2, s.sample(x * 0.25, steps=20, order=2))
    with pytest.raises(ValueError):
        s.capture(x, method="adaptive")


@pytest.mark.parametrize("name", ["pp2m", "pp3m", "eps3s_cfg", "pp2m_cfg_v", "eps2m_score", "pp3s_taylor"])
def test_prepared_steps_replay_equals_first_run(golden, cuda_backend, name):
    """Second and later sample() calls of a solver launch frozen descriptors (ops.PreparedStep: pointers patched,
    no per-step argument marshalling). They must reproduce the first (general-path) run and the reference golden
    bit for bit, follow a changed input, and fall back cleanly when the input looks different (other batch size)."""
    case = CASES[name]
    y1, _, _, s = run_product_case(case, device="cuda:0", return_solver=True)
    assert s._prep_cache, "no launch was frozen"
    x = __import__("cases").seeded(case["shape"], case["seed"]).cuda()
    kw = dict(steps=case["steps"], order=case["order"], skip_type=case["skip_type"], method=case["method"],
              lower_order_final=case.get("lower_order_final", True), denoise_to_zero=case.get("denoise_to_zero", False),
              solver_type=case.get("solver_type", "dpmsolver"), t_end=case.get("t_end"))
    before = cuda_backend.launch_count()
    y2 = s.sample(x, **kw)
    assert cuda_backend.launch_count() > before
    assert torch.equal(y2, y1)
    np.testing.assert_array_equal(y2.cpu().numpy(), golden["samples"][f"{name}/y"])
    y3 = s.sample(x * 0.5, **kw)                                 # new values, same descriptors
    s._prep_cache.clear()
    assert torch.equal(y3, s.sample(x * 0.5, **kw))             # == the general path
    if not case.get("cfg"):                                      # (the CFG cases carry per-sample conditions)
        y4 = s.sample(x[:1], **kw)                               # other batch size: other descriptors
        s._prep_cache.clear()
        assert torch.equal(y4, s.sample(x[:1], **kw))
