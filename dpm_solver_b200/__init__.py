"""dpm_solver_b200 -- B200-native (sm_100a) implementation of DPM-Solver's per-step update path.

Drop-in for the three public names of LuChengTHU/dpm-solver's `dpm_solver_pytorch.py`:

    from dpm_solver_b200 import NoiseScheduleVP, model_wrapper, DPM_Solver

The arithmetic runs in hand-written CUDA kernels behind a C-ABI shared library
(include/dpm_solver_b200.h, built by `python -m dpm_solver_b200.build`). There is no CPU or
PyTorch fallback: CUDA tensors only, and a missing library raises.
"""
from .schedule import NoiseScheduleVP, expand_dims, interpolate_fn
from .diffedit import DiffEditCorrector
from .solver import DPM_Solver, WrappedModel, model_wrapper

__version__ = "0.1.0"
__all__ = ["NoiseScheduleVP", "model_wrapper", "DPM_Solver", "WrappedModel", "interpolate_fn",
           "expand_dims", "DiffEditCorrector"]
