"""ctypes binding of librandblas_b200.so (the C ABI in include/randblas_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import of the symbols.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RANDBLAS_B200_LIB: an alternative build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("RANDBLAS_B200_LIB") or os.path.join(HERE, "librandblas_b200.so")

_lib = None


class RandBLASError(RuntimeError):
    """Mirror of RandBLAS::Error (reference: RandBLAS/exceptions.hh:57-95)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). randblas_b200 has no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.rb_last_error.restype = ctypes.c_char_p
        _lib.rb_get_counter.restype = ctypes.c_int64
    return _lib


_C = {"c": ctypes.c_char, "i": ctypes.c_int, "q": ctypes.c_int64, "Q": ctypes.c_uint64, "f": ctypes.c_float,
      "d": ctypes.c_double, "p": ctypes.c_void_p}


def call(name, sig, *args):
    """Call an int-returning C-ABI function; raise RandBLASError on a non-zero return."""
    fn = getattr(lib(), name)
    fn.restype = ctypes.c_int
    assert len(sig) == len(args), (name, len(sig), len(args))
    conv = []
    for code, a in zip(sig, args):
        if code == "p":
            if isinstance(a, (ctypes.Array, ctypes._SimpleCData)):
                conv.append(ctypes.byref(a))       # keeps the ctypes object alive for the duration of the call
            else:
                conv.append(ctypes.c_void_p(a) if a else None)
        elif code == "c":
            conv.append(ctypes.c_char(a.encode() if isinstance(a, str) else a))
        else:
            conv.append(_C[code](a))
    rc = fn(*conv)
    if rc != 0:
        raise RandBLASError(lib().rb_last_error().decode(errors="replace") + f" [rc={rc}]")
    return rc


def counter(name):
    return int(lib().rb_get_counter(name.encode()))


def get_option(name):
    fn = lib().rb_get_option
    fn.restype = ctypes.c_int64
    return int(fn(ctypes.c_char_p(name.encode())))


def set_option(name, value):
    fn = lib().rb_set_option
    fn.restype = ctypes.c_int
    if fn(ctypes.c_char_p(name.encode()), ctypes.c_int64(int(value))) != 0:
        raise RandBLASError(lib().rb_last_error().decode(errors="replace"))


def release_workspace():
    """rb_release_workspace: frees every cached scratch buffer of the library (no call may be in flight)."""
    fn = lib().rb_release_workspace
    fn.restype = ctypes.c_int
    if fn() != 0:
        raise RandBLASError(lib().rb_last_error().decode(errors="replace"))
