import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; see oracle/vxo.h)."""
    from oracle import vxo_py
    vxo_py.lib()
    return vxo_py


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from voxelengine_b200.build import build
    build()
    from voxelengine_b200.engine import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
