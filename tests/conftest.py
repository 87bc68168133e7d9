import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings("ignore", category=SyntaxWarning)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {n: np.load(os.path.join(GOLDEN, n + ".npz")) for n in ("schedules", "updates", "glue", "samples")}


@pytest.fixture()
def oracle_backend():
    """Run the product's host logic on the numpy executor (CPU tests only)."""
    from dpm_solver_b200 import ops
    from oracle_backend import OracleBackend
    be = OracleBackend()
    old = ops._backend
    ops.set_backend(be)
    yield be
    ops.set_backend(old)


@pytest.fixture()
def cuda_backend():
    from dpm_solver_b200 import ops
    old = ops._backend
    ops.set_backend(ops.CudaBackend())
    yield ops.backend()
    ops.set_backend(old)


@pytest.fixture()
def oracle_backend_cpu():
    from oracle_backend import OracleBackend
    return OracleBackend()
