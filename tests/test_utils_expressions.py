"""Host-side numeric utilities and expression algebra (reference: t/utils.lisp,
t/system-info.lisp, t/expressions.lisp) -- and the link to the device thresholds."""
from fractions import Fraction

import numpy as np
import pytest

from linear_programming_b200 import _ffi, conditions, expressions as E, utils as U

f32 = np.float32


def test_bound_helpers():
    """t/utils.lisp:18-58"""
    assert U.lb_min(None, 3) is None and U.lb_min(2, 3) == 2 and U.lb_min(2, None) is None
    assert U.lb_max(None, 3) == 3 and U.lb_max(2, 3) == 3 and U.lb_max(2, None) == 2
    assert U.ub_min(None, 3) == 3 and U.ub_min(2, 3) == 2 and U.ub_min(2, None) == 2
    assert U.ub_max(None, 3) is None and U.ub_max(2, 3) == 3 and U.ub_max(2, None) is None


def test_validate_bounds():
    """t/utils.lisp:60-70"""
    with pytest.raises(conditions.InvalidBoundsError):
        U.validate_bounds(5, -4, "x")
    for lb, ub in [(None, -4), (None, None), (5, None), (5, 6)]:
        U.validate_bounds(lb, ub, "x")


def test_optimization_type_and_contagion():
    """t/system-info.lisp"""
    assert U.optimization_type(3) == "rational" and U.optimization_type(Fraction(1, 3)) == "rational"
    assert U.optimization_type(f32(1)) == "single-float" and U.optimization_type(1.0) == "double-float"
    assert U.float_contagion("rational", "rational") == "rational"
    assert U.float_contagion("rational", "single-float") == "single-float"
    assert U.float_contagion("double-float", "rational") == "double-float"
    assert U.float_contagion("single-float", "double-float") == "double-float"


@pytest.mark.parametrize("zero,eps", [(f32(0), U.SINGLE_FLOAT_EPSILON), (0.0, U.DOUBLE_FLOAT_EPSILON)])
def test_fp_predicates(zero, eps):
    """t/utils.lisp:73-157"""
    kind = type(zero)
    assert U.fp_eq(0, 0) and U.fp_eq(8, 8) and not U.fp_eq(0, 1) and not U.fp_eq(0, Fraction(1, 2 ** 128))
    assert U.fp_eq(zero, zero) and U.fp_eq(zero, kind(4 * eps)) and not U.fp_eq(zero, kind(4 * eps), 1)
    assert not U.fp_eq(kind(0.01), zero)
    assert U.fp_gt(0, -1) and U.fp_gt(Fraction(1, 2 ** 128), 0) and not U.fp_gt(0, 0)
    assert not U.fp_ge(0, Fraction(1, 2 ** 128))
    assert U.fp_gt(zero, kind(-1)) and not U.fp_gt(zero, zero)
    assert not U.fp_gt(kind(4 * eps), zero) and U.fp_gt(kind(4 * eps), zero, 1)
    assert U.fp_lt(-1, 0) and U.fp_lt(0, Fraction(1, 2 ** 128)) and not U.fp_lt(0, 0)
    assert U.fp_lt(kind(-1), zero) and not U.fp_lt(zero, zero)
    assert not U.fp_lt(zero, kind(4 * eps)) and U.fp_lt(zero, kind(4 * eps), 1)
    assert U.fp_le(zero, zero) and U.fp_ge(zero, zero) and U.fp_le(kind(4 * eps), zero)


def test_device_thresholds_are_the_host_predicates():
    """What k_iter's look role compares against (b200lp_thresholds) == fp< / fp> / fp= with the
    reference's factors: tol/8 entering (src/simplex.lisp:370-378), tol/2 pivot element
    (:386-387), tol feasibility (:405-406)."""
    for tol in (1024.0, 16.0, 1.0):
        enter, pivot, feas = _ffi.thresholds(tol)
        for thr, factor in [(enter, tol / 8), (pivot, tol / 2)]:
            below, above = np.nextafter(thr, 0.0), np.nextafter(thr, 1.0)
            assert not U.fp_lt(0, float(below), factor) and not U.fp_lt(0, float(thr), factor)
            assert U.fp_lt(0, float(above), factor)
            assert U.fp_gt(float(above), 0, factor) and not U.fp_gt(float(thr), 0, factor)
            assert U.fp_lt(-float(above), 0, factor) and not U.fp_lt(-float(thr), 0, factor)
        assert U.fp_eq(0, float(feas), tol) and not U.fp_eq(0, float(np.nextafter(feas, 1.0)), tol)


def test_linear_expression_algebra():
    """t/expressions.lisp"""
    C = E.CONSTANT
    assert E.sum_linear_expressions({"a": 1, "b": 2}, {"a": 3, "c": 5}, {"b": 2}) == dict(a=4, b=4, c=5)
    assert E.scale_linear_expression({"a": 1, "b": 2}, 3) == dict(a=3, b=6)
    assert E.parse_linear_expression("x") == {"x": 1} and E.parse_linear_expression(5) == {C: 5}
    assert E.parse_linear_expression(["+", "x", ["*", 2, "y"], 3]) == {"x": 1, "y": 2, C: 3}
    assert E.parse_linear_expression(["*", 2, ["+", "x", 1], 3]) == {"x": 6, C: 6}
    assert E.parse_linear_expression(["-", "x"]) == {"x": -1}
    assert E.parse_linear_expression(["-", "x", "y", 2]) == {"x": 1, "y": -1, C: -2}
    assert E.parse_linear_expression(["/", "x", 2, 2]) == {"x": Fraction(1, 4)}
    assert E.parse_linear_expression(["/", 4]) == {C: Fraction(1, 4)}
    assert E.parse_linear_expression([":alist", ["x", 2], ["y", 3]]) == {"x": 2, "y": 3}
    assert E.parse_linear_expression([":plist", "x", 2, "y", 3]) == {"x": 2, "y": 3}
    for bad in (["*", "x", "y"], ["/", 2, "x"], ["/", "x"], ["expt", "x", 2]):
        with pytest.raises(conditions.NonlinearError):
            E.parse_linear_expression(bad)
    with pytest.raises(conditions.ParsingError):
        E.parse_linear_expression(None)
    assert E.format_linear_expression({"x": 2, C: 3}) == ["+", ["*", 2, "x"], 3]
