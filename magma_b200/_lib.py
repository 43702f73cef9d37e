"""ctypes loader for libmagma_b200.so (the C ABI declared in include/magma_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100 every
compute entry point raises. PyTorch is used only for device memory and streams.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmagma_b200.so")

_lib = None


class MB200Error(RuntimeError):
    pass


class Operand(ctypes.Structure):
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("ld", ctypes.c_int64),
        ("bs0", ctypes.c_int64),
        ("bs1", ctypes.c_int64),
        ("mn_major", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
    ]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32),
        ("N", ctypes.c_int32),
        ("K", ctypes.c_int32),
        ("nb0", ctypes.c_int32),
        ("nb1", ctypes.c_int32),
        ("c_dtype", ctypes.c_int32),
        ("A", Operand),
        ("B", Operand),
        ("C", ctypes.c_void_p),
        ("ldc", ctypes.c_int64),
        ("c_bs0", ctypes.c_int64),
        ("c_bs1", ctypes.c_int64),
        ("alpha", ctypes.c_float),
        ("act", ctypes.c_int32),
        ("dact", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("aux_out", ctypes.c_void_p),
        ("aux_in", ctypes.c_void_p),
        ("res1", ctypes.c_void_p),
        ("res2", ctypes.c_void_p),
        ("ld_res", ctypes.c_int64),
        ("force_bn", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
    ]


def lib():
    """Load (once) and return the ctypes handle. Raises MB200Error when the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MB200Error(
                f"{LIB_PATH} not found: build it with `python -m magma_b200.build` "
                "(magma_b200 has no CPU / eager fallback)"
            )
        L = ctypes.CDLL(LIB_PATH)
        L.mb200_last_error.restype = ctypes.c_char_p
        L.mb200_version.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MB200Error(f"magma_b200 error {rc}: {lib().mb200_last_error().decode()}")
