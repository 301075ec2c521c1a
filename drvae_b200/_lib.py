"""ctypes binding of the C ABI declared in include/drvae_b200.h.

There is no CPU fallback: if the shared library has not been built (python -m drvae_b200.build)
or a call fails, a RuntimeError is raised with the library's own message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdrvae_b200.so")
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "drvae_b200: %s is missing. Build it with `python -m drvae_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.drvae_last_error.restype = ctypes.c_char_p
    lib.drvae_debug_gemm.restype = c_int
    lib.drvae_debug_gemm.argtypes = [
        c_int, c_int, c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_ll,
        c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().drvae_last_error()
        raise RuntimeError("drvae_b200 %s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))
