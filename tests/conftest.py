import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the oracle's C++ part are built once per session (nvcc/g++ cross-compile on CPU)."""
    import __graft_entry__ as g
    g.build()


def small_flags(**kw):
    from uresnet_pytorch_b200.model import Flags
    base = dict(DATA_DIM=3, SPATIAL_SIZE=32, URESNET_FILTERS=8, URESNET_NUM_STRIDES=3, NUM_CLASS=5)
    base.update(kw)
    return Flags(**base)
