import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sparse_lists(idx, dist, cnt):
    import numpy as np
    m = np.arange(idx.shape[1])[None, :] < cnt[:, None]
    return idx[m], (dist[m] if dist is not None else None)


@pytest.fixture(scope="session")
def golden_default():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "default_scene.npz"))


@pytest.fixture(scope="session")
def golden_full():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "full_dambreak_16k.npz"))
