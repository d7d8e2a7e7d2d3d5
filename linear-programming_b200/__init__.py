"""linear_programming_b200 -- B200-native dense simplex backend for the `*solver*` hook of
neil-lindquist/linear-programming (reference @ 7fe5c78).

csrc/         CUDA kernels + the C ABI (include/b200lp.h) -> libb200lp.so
_ffi.py       ctypes binding of that ABI (what the Lisp CFFI shim binds too)
lisp/         the CFFI shim installing `b200-solver` into `*solver*`
"""
from . import _ffi  # noqa: F401
from ._ffi import (B200DeviceError, B200LibraryError, DeviceTableau, make_opts)  # noqa: F401

__all__ = ["_ffi", "B200DeviceError", "B200LibraryError", "DeviceTableau", "make_opts"]
