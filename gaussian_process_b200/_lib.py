"""ctypes binding of libgpx.so (the C ABI declared in include/gpx.h).

There is NO CPU fallback: if the shared library is missing or no B200 is visible, every compute
entry point raises.  The prototypes below are parsed from ``include/gpx.h`` so the binding cannot
drift from the header.
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libgpx.so")
HEADER_PATH = os.path.join(_ROOT, "include", "gpx.h")

GPX_TILE = 128
COV_SE, COV_LIN, COV_PER, COV_CO2 = 0, 1, 2, 3
COV_SAME_X, COV_LOWER, COV_DELTA = 1, 2, 4


class GpxError(RuntimeError):
    pass


_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "long long": ctypes.c_longlong,
    "double": ctypes.c_double,
    "void": None,
    "const char*": ctypes.c_char_p,
    "gpx_handle": ctypes.c_void_p,
    "gpx_handle*": ctypes.POINTER(ctypes.c_void_p),
}


def _ctype_of(decl: str):
    decl = decl.strip()
    decl = re.sub(r"\s+", " ", decl)
    # drop the parameter name
    decl = decl.replace("long long", "longlong")
    m = re.match(r"^(const )?(\w+)\s*(\*?)\s*(\w+)?$", decl)
    if not m:
        raise ValueError("cannot parse parameter %r" % decl)
    const, base, star, _name = m.groups()
    if star:
        if base == "gpx_handle":
            return ctypes.POINTER(ctypes.c_void_p)
        return ctypes.c_void_p  # every data pointer crosses as a raw address
    return _CTYPES[base]


def parse_header(path: str = HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every function declared in gpx.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*(int64_t|int|const char\*|void)\s+(gpx_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = args.strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (_CTYPES[ret], argtypes)
    return protos


_lib = None


def load():
    """Load libgpx.so (once) and attach prototypes.  Raises GpxError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise GpxError(
            "libgpx.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().gpx_last_error().decode("utf-8", "replace")


def check(status: int, what: str = "gpx"):
    """Map a libgpx status to the exceptions the reference raises."""
    if status == 0:
        return
    if status > 0:
        # np.linalg.cholesky's failure mode (GP_regression.py:154 etc.)
        raise np.linalg.LinAlgError("Matrix is not positive definite (%s: leading minor of order %d)" % (what, status))
    raise GpxError("%s failed with status %d: %s" % (what, status, last_error()))
