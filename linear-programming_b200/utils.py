"""Bound helpers and tolerance predicates (reference: src/utils.lisp, src/system-info.lisp).

The device side compares with the same absolute thresholds: `factor * double-float-epsilon`
(b200lp_thresholds in include/b200lp.h) -- fp_lt / fp_gt / fp_eq here are the host-side statement
of what k_iter's look role tests with factor = tol/8 (entering column, src/simplex.lisp:370-378),
tol/2 (pivot element, :386-387) and tol (phase-1 feasibility, :405-406).

Numeric kinds: int / Fraction are CL rationals (compared exactly), numpy.float32 is single-float,
Python float / numpy.float64 is double-float; contagion picks the wider float (system-info.lisp:38-63).
"""
from fractions import Fraction
from numbers import Rational

import numpy as np

from .conditions import InvalidBoundsError

# CL epsilons as SBCL defines them: the smallest e with (/= 1 (+ 1 e)), i.e. 2^-p (1 + 2^-(p-1))
SINGLE_FLOAT_EPSILON = float.fromhex("0x1.000002p-24")        # 5.960465e-8
DOUBLE_FLOAT_EPSILON = float.fromhex("0x1.0000000000001p-53")  # 1.1102230246251568e-16


def lb_min(x, y):
    """src/utils.lisp:36-42 -- None is negative infinity"""
    return None if x is None or y is None else min(x, y)


def lb_max(x, y):
    """:44-50"""
    return y if x is None else x if y is None else max(x, y)


def ub_min(x, y):
    """:52-58 -- None is positive infinity"""
    return y if x is None else x if y is None else min(x, y)


def ub_max(x, y):
    """:60-66"""
    return None if x is None or y is None else max(x, y)


def validate_bounds(lb, ub, var):
    """:69-76"""
    if lb is not None and ub is not None and ub < lb:
        raise InvalidBoundsError(var, ub, lb)


def optimization_type(x):
    """src/system-info.lisp:31-36"""
    if isinstance(x, (Rational, Fraction)) and not isinstance(x, bool):
        return "rational"
    if isinstance(x, np.float32):
        return "single-float"
    if isinstance(x, (float, np.floating)):
        return "double-float"
    raise TypeError(f"{x!r} is not a real number")


def float_contagion(t1, t2):
    """src/system-info.lisp:38-63 with the two float types of this port"""
    if t1 == t2:
        return t1
    if t1 == "rational":
        return t2
    if t2 == "rational":
        return t1
    return "double-float"


def _eps(a, b):
    kind = float_contagion(optimization_type(a), optimization_type(b))
    return {"rational": None, "single-float": SINGLE_FLOAT_EPSILON,
            "double-float": DOUBLE_FLOAT_EPSILON}[kind]


def fp_eq(a, b, factor=16):
    """fp=, src/utils.lisp:84-93"""
    e = _eps(a, b)
    return a == b if e is None else abs(float(a) - float(b)) <= factor * e


def fp_le(a, b, factor=16):
    """fp<=, :121"""
    e = _eps(a, b)
    return a <= b if e is None else float(a) <= float(b) + factor * e


def fp_ge(a, b, factor=16):
    """fp>=, :122"""
    e = _eps(a, b)
    return a >= b if e is None else float(a) >= float(b) - factor * e


def fp_lt(a, b, factor=16):
    """fp<, :123 -- a < b - factor*eps"""
    e = _eps(a, b)
    return a < b if e is None else float(a) < float(b) - factor * e


def fp_gt(a, b, factor=16):
    """fp>, :124 -- a > b + factor*eps"""
    e = _eps(a, b)
    return a > b if e is None else float(a) > float(b) + factor * e


__all__ = ["lb_min", "lb_max", "ub_min", "ub_max", "validate_bounds", "optimization_type",
           "float_contagion", "fp_eq", "fp_le", "fp_ge", "fp_lt", "fp_gt",
           "SINGLE_FLOAT_EPSILON", "DOUBLE_FLOAT_EPSILON"]
