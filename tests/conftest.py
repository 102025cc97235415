import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_forward():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "forward.npz"))


@pytest.fixture(scope="session")
def golden_inverse():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "inverse.npz"))


@pytest.fixture(scope="session")
def golden_minv():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "matrix_inverse.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) the in-tree C-ABI library; nvcc cross-compiles without a GPU."""
    from hdr2yuv_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(built_lib):
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from hdr2yuv_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture
def opt(ctx):
    """Set test switches of the shared context (h2y_ctx_set_option); the defaults come back after the test."""
    def set_option(name, value):
        ctx.set_option(name, value)
    yield set_option
    ctx.set_option(None)
