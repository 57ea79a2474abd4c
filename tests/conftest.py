import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cli-p_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_c():
    """The plain-C oracle (oracle/flatip_ref.c), built on demand."""
    import ctypes as C
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_flatip.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(so)
    lib.oracle_flatip_search.restype = C.c_int
    lib.oracle_flatip_search.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_int]

    import numpy as np

    def search(xb, xq, k, threads=0):
        xb = np.ascontiguousarray(xb)
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        dtype = 1 if xb.dtype == np.float16 else 0
        assert xb.dtype in (np.float16, np.float32)
        nq = xq.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        rc = lib.oracle_flatip_search(xb.ctypes.data, dtype, xb.shape[0], xb.shape[1], xq.ctypes.data, nq, k,
                                      D.ctypes.data, I.ctypes.data, threads)
        assert rc == 0
        return D, I

    return search
