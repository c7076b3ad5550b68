import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


# the CPU oracle on one thread: fixed summation order, bit-reproducible on any machine (oracle/orc_ba.c: orc_apply_threading)
os.environ.setdefault("ORC_DETERMINISTIC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def mm():
    """The product package with the CUDA library loaded (fails loudly if it is missing)."""
    import mavmap_b200
    from mavmap_b200 import _lib
    _lib.lib()
    if _lib.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible")
    return mavmap_b200
