"""The oracle itself against the UNMODIFIED reference on random sampling configurations (build container only):
torch-CPU namespace bit-identical, numpy namespace within its transcendental-ulp tolerance. Complements the
fixed golden vectors that pin the oracle (tests/test_oracle_golden.py)."""
import random

import numpy as np
import pytest
import torch

import helpers
from oracle import dpm_oracle as O
from test_random_configs_vs_reference import draw, pytestmark, reference_module, run  # noqa: F401  (same skip rule)


@pytest.mark.parametrize("chunk", range(3))
def test_oracle_matches_reference_on_random_configurations(chunk):
    ref = reference_module()
    rng = random.Random(70000 + chunk)
    done = 0
    while done < 20:
        c = draw(rng)
        case = dict(schedule=c["schedule"], algo=c["algo"], method=c["method"], order=c["order"], steps=c["steps"],
                    skip_type=c["skip_type"], solver_type=c["solver_type"], model_type=c["model_type"], cfg=c["cfg"],
                    lower_order_final=c["lower_order_final"], denoise_to_zero=c["denoise_to_zero"], t_end=c["t_end"],
                    seed=c["seed"], thresholding=c["thresholding"], shape=(2, 3, 8, 8), net="exact")
        try:
            yr, _, _ = run(ref.NoiseScheduleVP, ref.model_wrapper, ref.DPM_Solver, c)
        except Exception:
            continue
        if not torch.isfinite(yr).all():
            continue
        yt, _, _ = helpers.run_oracle_case(case, None, O.torch_namespace("cpu"))
        yt = yt.numpy() if torch.is_tensor(yt) else np.asarray(yt)
        np.testing.assert_array_equal(yt, yr.numpy(), err_msg=str(case))
        yn, _, _ = helpers.run_oracle_case(case, None, O.NP)
        yn = yn.numpy() if torch.is_tensor(yn) else np.asarray(yn)
        assert helpers.rel_err(yn, yr.numpy()) <= 1e-3, case
        done += 1
