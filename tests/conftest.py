import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _cuda_device_count():
    try:
        import ctypes as C
        from voronoirt_b200 import _lib
        n = C.c_int32(0)
        return n.value if _lib.lib().vrt_device_count(C.byref(n)) == 0 else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device, whatever -m says"""
    if not any("gpu" in it.keywords for it in items) or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libvrt has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_grid(name):
    """committed fixture written by tests/make_golden.py: positions (3,n) rows (z,x,y), NeighbourMatrix (n,ld), bounds"""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    pos = np.asfortranarray(d["positions"], dtype=np.float64)
    nbr = np.asfortranarray(d["neighbours"].astype(np.int64))
    return pos, nbr, d["bounds"].astype(np.float64)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


def oracle_sites(O, pos, nbr, bounds):
    """oracle Sites from the Fortran-ordered (Julia-shaped) arrays"""
    return O.Sites(np.ascontiguousarray(pos.T), np.ascontiguousarray(nbr.T), bounds)
