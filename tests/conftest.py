import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle .so files exist (gcc only; the _ref build needs /root/reference and is
    otherwise taken prebuilt) and that the product library is built."""
    from oracle import oracle as o
    o.build()
    from spacefortress_b200 import build as b
    if os.environ.get("SF_SKIP_NVCC_BUILD") != "1":
        b.build()
    yield


def scripted_kill_policy(t, vulnerability):
    """Autoturn key masks that raise the vulnerability with slow shots (> 250 ms apart), then double-tap.
    Returns (FIRE bit) pattern: slow phase = press every 10th tick; kill phase = press on alternate ticks."""
    if vulnerability < 11:
        return 1 if (t % 10) == 0 else 0
    return 1 if (t % 2) == 0 else 0
