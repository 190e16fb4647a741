import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import flexq_oracle
    return flexq_oracle


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle on the plain-C restatement (built on demand with gcc)."""
    import ctypes
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "libflexq_oracle.so")
    src = os.path.join(ROOT, "oracle", "flexq_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libflexq_oracle.so"])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
