"""ctypes front end of the fp64 C oracle (oracle/simplex_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE -- only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_simplex.so")

OPTIMAL, UNBOUNDED, INFEASIBLE, ITERATION_LIMIT, ARTIFICIAL_STUCK, ARTIFICIAL_NONZERO = 0, 1, 2, 3, 4, 5
FEAS_SCALED, FEAS_REFERENCE = 0, 1

_lib = None
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "simplex_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_simplex.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_cl_epsilon.restype = ctypes.c_double
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_find_entering_column.restype = ctypes.c_int64
        L.oracle_find_entering_column.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.oracle_find_pivoting_row.restype = ctypes.c_int64
        L.oracle_find_pivoting_row.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            ctypes.c_int64, ctypes.c_double, ctypes.c_int]
        L.oracle_pivot.restype = None
        L.oracle_pivot.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.oracle_solve.restype = ctypes.c_int
        L.oracle_solve.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
            _c_int64_p, _c_int32_p, _c_int32_p, ctypes.c_int64]
        L.oracle_solve_cycle_probe.restype = ctypes.c_int
        L.oracle_solve_cycle_probe.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _c_int64_p]
        L.oracle_solve_two_phase.restype = ctypes.c_int
        L.oracle_solve_two_phase.argtypes = [
            _c_double_p, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_int32_p,
            ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, _c_int64_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int32_p)


def _check_tab(tab):
    assert tab.dtype == np.float64 and tab.ndim == 2 and tab.flags.c_contiguous
    return tab.shape[0], tab.shape[1], tab.strides[0] // 8


def cl_epsilon():
    return lib().oracle_cl_epsilon()


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    """All host cores for the CPU-baseline legs (torchrun pins OMP_NUM_THREADS=1)."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def find_entering_column(tab, is_max, tol=1024.0, rule=0):
    R, C, ld = _check_tab(tab)
    return int(lib().oracle_find_entering_column(_dp(tab), R, C, ld, int(is_max), tol, rule))


def find_pivoting_row(tab, basis, j, tol=1024.0, rule=0):
    R, C, ld = _check_tab(tab)
    return int(lib().oracle_find_pivoting_row(_dp(tab), R, C, ld, _ip(basis), j, tol, rule))


def pivot(tab, basis, j, p, parallel=False):
    """In place, like n-pivot-row."""
    R, C, ld = _check_tab(tab)
    lib().oracle_pivot(_dp(tab), R, C, ld, _ip(basis), j, p, int(parallel))


def solve(tab, basis, is_max=True, tol=1024.0, rule=0, max_iters=0, parallel=False,
          trace_cap=0):
    """In place, like n-solve-tableau. Returns (status, iterations, trace[(j, r)...])."""
    R, C, ld = _check_tab(tab)
    assert basis.dtype == np.int32 and basis.shape == (R - 1,)
    iters = ctypes.c_int64(0)
    tj = np.zeros(max(trace_cap, 1), np.int32)
    tr = np.zeros(max(trace_cap, 1), np.int32)
    st = lib().oracle_solve(_dp(tab), R, C, ld, _ip(basis), int(is_max), tol, rule,
                            max_iters, int(parallel), ctypes.byref(iters),
                            _ip(tj), _ip(tr), trace_cap)
    n = min(iters.value, trace_cap)
    return st, iters.value, list(zip(tj[:n].tolist(), tr[:n].tolist()))


def solve_cycle_probe(tab, basis, is_max=True, tol=1024.0, rule=0, max_iters=100000, parallel=False):
    """n-solve-tableau in place with a basis-revisit detector (a revisited basis under a
    deterministic rule is a proven cycle).  Returns (status, dict(pivots, degenerate_pivots,
    revisit_at, first_visit_at)); revisit_at is -1 when no basis was seen twice."""
    R, C, ld = _check_tab(tab)
    out = (ctypes.c_int64 * 4)()
    st = lib().oracle_solve_cycle_probe(_dp(tab), R, C, ld, _ip(basis), int(is_max), tol, rule,
                                        max_iters, int(parallel), out)
    return st, dict(pivots=out[0], degenerate_pivots=out[1], revisit_at=out[2], first_visit_at=out[3])


def solve_two_phase(art, art_basis, main, main_basis, is_max=True, tol=1024.0, rule=0,
                    max_iters=0, parallel=False, feas_mode=FEAS_SCALED, with_redundant=False):
    """In place on both tableaus. Returns (status, (phase1, cleanup, phase2) pivots)
    [+ redundant rows when with_redundant]."""
    R, C_art, ld_art = _check_tab(art)
    R2, C, ld = _check_tab(main)
    assert R == R2
    iters = (ctypes.c_int64 * 4)()
    st = lib().oracle_solve_two_phase(_dp(art), C_art, ld_art, _ip(art_basis),
                                      _dp(main), R, C, ld, _ip(main_basis),
                                      int(is_max), tol, rule, max_iters, int(parallel),
                                      int(feas_mode), iters)
    return st, (tuple(iters) if with_redundant else tuple(iters)[:3])
