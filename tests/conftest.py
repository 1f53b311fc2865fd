import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
        return {k: z[k] for k in z.files}
    return load


@pytest.fixture(scope="session")
def host_shim():
    """g++ build of tests/host_shim.cpp: frb_math.h / frb_head.h (the kernels' arithmetic) on the CPU."""
    import ctypes
    so = os.path.join(ROOT, "tests", "_host_shim.so")
    src = os.path.join(ROOT, "tests", "host_shim.cpp")
    hdrs = [os.path.join(ROOT, "fresnel_b200", "csrc", h) for h in ("frb_math.h", "frb_head.h")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def cuda_lib():
    from fresnel_b200 import _lib
    return _lib.lib()
