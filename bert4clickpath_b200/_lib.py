"""ctypes binding of libb4cp.so — the only door from Python into the CUDA kernels.

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb4cp.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b4cp.h")

_lib = None


class B4cpError(RuntimeError):
    pass


class GemmEpilogue(ctypes.Structure):
    """Mirror of b4cp_gemm_epilogue (include/b4cp.h)."""
    _fields_ = [
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p),
        ("relu", ctypes.c_int),
        ("gate", ctypes.c_void_p),
        ("ld_gate", ctypes.c_long),
        ("addend", ctypes.c_void_p),
        ("ld_addend", ctypes.c_long),
        ("out_f32", ctypes.c_void_p),
        ("ld_f32", ctypes.c_long),
        ("split_stride", ctypes.c_long),
        ("out_bf16", ctypes.c_void_p),
        ("ld_bf16", ctypes.c_long),
    ]


class AdamSegment(ctypes.Structure):
    """Mirror of b4cp_adam_segment (include/b4cp.h)."""
    _fields_ = [
        ("begin", ctypes.c_long),
        ("numel", ctypes.c_long),
        ("cols", ctypes.c_int),
        ("ld_shadow", ctypes.c_long),
        ("shadow_bf16", ctypes.c_void_p),
    ]


ADAM_MAX_SEGS = 48


def declared_symbols():
    """Every function name declared in include/b4cp.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b4cp_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load libb4cp.so once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B4cpError(
                f"{LIB_PATH} not found: build it with `python -m bert4clickpath_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b4cp_last_error.restype = ctypes.c_char_p
    return _lib


def call(name, *args):
    """Invoke an int-returning export and raise with b4cp_last_error() on failure."""
    L = lib()
    fn = getattr(L, name)
    fn.restype = ctypes.c_int
    rc = fn(*args)
    if rc != 0:
        raise B4cpError(f"{name} failed (rc={rc}): {L.b4cp_last_error().decode()}")
    return rc


# ---- argument helpers -------------------------------------------------------------------
def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def c_int(v):
    return ctypes.c_int(int(v))


def c_long(v):
    return ctypes.c_long(int(v))


def c_float(v):
    return ctypes.c_float(float(v))


def c_u64(v):
    return ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
