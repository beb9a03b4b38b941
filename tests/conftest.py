import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _ensure_built():
    need = [os.path.join(ROOT, "sahara_b200", "libsahara_b200.so"), os.path.join(ROOT, "sahara_b200", "libsahara_host.so"),
            os.path.join(ROOT, "oracle", "libsahara_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "all"], stdout=subprocess.DEVNULL)


_ensure_built()


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(12345)
