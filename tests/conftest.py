import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_pod():
    return np.load(os.path.join(GOLDEN, "pod_from_data_ref.npz"))


@pytest.fixture(scope="session")
def golden_jtj():
    return np.load(os.path.join(GOLDEN, "meanjtj_ref.npz"))


@pytest.fixture(scope="session")
def golden_dp():
    return np.load(os.path.join(GOLDEN, "doublepass_ref.npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def subspace_angle(U, V, M=None):
    from oracle.projectors_np import principal_angle
    return principal_angle(U, V, M)
