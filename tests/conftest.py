import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_stage1():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_stage1.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_octaves():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_octaves.npz"), allow_pickle=False)
