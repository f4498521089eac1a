"""ctypes loader for libpvcr_b200.so (the C ABI in include/pvcr_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvcr_b200.so")

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_size = ctypes.c_size_t
c_vp = ctypes.c_void_p
c_f = ctypes.c_float
c_u64 = ctypes.c_uint64

_lib = None


class PvcrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvcrError("%s not found: build it with `python __graft_entry__.py` (no CPU fallback exists)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pvcr_last_error.restype = ctypes.c_char_p
        _lib.pvcr_version.restype = c_int
        _declare(_lib)
    return _lib


def _sig(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def _declare(L):
    _sig(L.pvcr_linear_fwd_workspace, c_size, [c_int, c_int, c_int, c_int])
    _sig(L.pvcr_linear_fwd, c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp,
                                    c_size, c_vp])
    _sig(L.pvcr_linear_bwd_workspace, c_size, [c_int, c_int, c_int, c_int])
    _sig(L.pvcr_linear_bwd, c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int,
                                    c_int, c_int, c_int, c_int, c_vp, c_size, c_vp])


def check(rc, what):
    if rc != 0:
        raise PvcrError("%s failed (%d): %s" % (what, rc, lib().pvcr_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
