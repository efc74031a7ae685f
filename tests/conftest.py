import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_numpy():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "numpy_half.npz"))


@pytest.fixture(scope="session")
def golden_tf():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "tf_half.npz"))


@pytest.fixture(scope="session")
def rn():
    """The product package with its CUDA library built and loaded (fails loudly otherwise)."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import retinanet_b200
    retinanet_b200._lib.load()
    return retinanet_b200
