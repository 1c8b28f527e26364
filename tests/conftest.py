import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the tensor-core attention kernel normally takes >= 1024 keys (where it wins); the tests push every shape it CAN take
# (>= 64 keys) through it, so the small networks of the parity tests exercise it too
os.environ.setdefault("GG_ATTN_TC_MIN_TK", "64")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def have_reference():
    from oracle import refshim
    return refshim.available()
