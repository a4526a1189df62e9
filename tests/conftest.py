import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    from gonova_tts_b200 import build, _cabi

    build.build()
    return _cabi.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("test is marked gpu but no CUDA device is visible")
    return torch.device("cuda:0")


def snr_db(test, ref):
    import numpy as np

    test = np.asarray(test, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    num = float((ref ** 2).sum())
    den = float(((test - ref) ** 2).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(num / den)
