"""Loader of the in-tree C-ABI library voronoirt_b200/libvrt.so.

There is no CPU fallback and no alternate backend: if the library is missing (not built) or cannot be
loaded, importing the compute API raises.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvrt.so")
_LIB = None


class VRTError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvrt error {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile voronoirt_b200/csrc/*.cu into voronoirt_b200/libvrt.so (sm_100a, -lineinfo)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libvrt.so failed")
    return LIB_PATH


def lib():
    """The bound library handle.  Fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "voronoirt_b200 has no CPU fallback.")
        _LIB = _abi.bind(C.CDLL(LIB_PATH))
    return _LIB


def check(rc):
    if rc != 0:
        raise VRTError(rc, lib().vrt_last_error().decode(errors="replace"))


def last_stats():
    out = (C.c_double * 8)()
    check(lib().vrt_last_stats(out))
    return {"kernels": out[0], "visits": out[1], "steps": out[2], "sweep_ms": out[3]}
