"""read-sexp / write-sexp / read-mps / write-standard-format (reference: t/external-formats.lisp).
CPU only; the last test runs an MPS problem end to end on the GPU."""
import io
import os
from fractions import Fraction

import pytest

from linear_programming_b200 import conditions, external_formats as X, problem as P, solver

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mps")


def norm(constraints):
    return {(op, tuple(sorted(terms)), rhs) for op, terms, rhs in constraints}


SIMPLE = """((max (+ x (* 4 y) (* 8 z)))
             (<= (+ x y) 8)
             (<= (+ (* 2 y) z) 7))"""


def test_read_sexp():
    """t/external-formats.lisp:23-41"""
    p = X.read_sexp(io.StringIO(SIMPLE))
    assert isinstance(p, P.Problem) and p.type == "max" and isinstance(p.objective_var, P.Uninterned)
    assert set(p.vars) == {"x", "y", "z"} and dict(p.objective_func) == dict(x=1, y=4, z=8)
    assert not p.integer_vars and not p.var_bounds
    assert norm(p.constraints) == {("<=", (("x", 1), ("y", 1)), 8), ("<=", (("y", 2), ("z", 1)), 7)}


def test_read_sexp_package_qualified_symbols_and_bounds():
    """t/external-formats.lisp:43-60"""
    p = X.read_sexp("""((max (+ x (* 4 y) (* 8 z))) (<= (+ x y) 8) (<= (+ y z) 7)
                        (linear-programming/problem:bounds (y)))""")
    assert p.var_bounds == [("y", (None, None))]


def test_read_sexp_read_eval_is_off_by_default():
    """t/external-formats.lisp:62-82"""
    text = "((max (+ x (* 4 y) (* 8 z))) (<= (+ x y) #.(+ 4 4)) (<= (+ y z) 7))"
    with pytest.raises(conditions.ParsingError):
        X.read_sexp(text)
    p = X.read_sexp(text, allow_read_eval=True)
    assert norm(p.constraints) == {("<=", (("x", 1), ("y", 1)), 8), ("<=", (("y", 1), ("z", 1)), 7)}


def test_read_sexp_consumes_only_one_form():
    """t/external-formats.lisp:104-122"""
    stream = io.StringIO(SIMPLE + "456")
    p = X.read_sexp(stream)
    assert stream.read() == "456" and set(p.vars) == {"x", "y", "z"}


@pytest.mark.parametrize("objective,constraints", [
    ("(max (+ x (* 4 y) (* 8 z)))", ["(<= (+ x y) 8)", "(<= (+ y z) 7)"]),
    ("(max (+ x (* 4 y) (* 8 z)))", ["(<= (+ x y) 8)", "(<= (+ y z) 7)", "(bounds (y))"]),
    ("(min (= w (+ (* 0.2d0 x) y)))", ["(>= (+ x y) 4.2d0)", "(integer x)"]),
    ("(max (+ x y))", ["(<= (+ (* 1/3 x) y) 5/2)", "(bounds (1 x 4) (y 9) (-2 z))", "(binary b)",
                       "(<= (+ x y z b) 10)"]),
])
def test_write_sexp_round_trips(objective, constraints):
    """t/external-formats.lisp:125-214"""
    base = P.make_linear_problem(objective, *constraints)
    out = io.StringIO()
    X.write_sexp(out, base)
    back = X.read_sexp(out.getvalue())
    assert back.type == base.type and set(back.vars) == set(base.vars)
    assert dict(back.objective_func) == dict(base.objective_func)
    assert back.integer_vars == base.integer_vars and dict(back.var_bounds) == dict(base.var_bounds)
    assert norm(back.constraints) == norm(base.constraints)
    named = not isinstance(base.objective_var, P.Uninterned)
    assert (back.objective_var == base.objective_var) if named else isinstance(back.objective_var, P.Uninterned)


@pytest.mark.parametrize("name", ["simple-problem.mps", "simple-problem-crlf.mps"])
def test_read_mps_simple(name):
    """t/external-formats.lisp:216-233, 289-306 (LF and CRLF line endings)"""
    with open(os.path.join(DATA, name), newline="") as f:
        p = X.read_mps(f, "max")
    assert p.type == "max" and isinstance(p.objective_var, P.Uninterned)
    assert set(p.vars) == {"x", "y", "z"} and dict(p.objective_func) == dict(x=1, y=4, z=8)
    assert not p.integer_vars and not p.var_bounds
    assert norm(p.constraints) == {("<=", (("x", 3), ("y", 1)), 8), ("<=", (("y", 1), ("z", 2)), 7)}


def test_read_mps_advanced():
    """t/external-formats.lisp:234-252: objsense, second RHS vector, negative RHS on a G row,
    BV/LO/UP/FR bounds, exact rational coefficients, ENDATA ends the read."""
    with open(os.path.join(DATA, "advanced-problem.mps")) as f:
        p = X.read_mps(f, None, read_case="preserve", rhs_id="rhs1")
        assert f.readline().startswith("trailing text")
    assert p.type == "min" and set(p.vars) == {"w", "X", "Y", "Z"}
    assert dict(p.objective_func) == {"w": -1, "X": 1, "Y": Fraction(9, 2), "Z": 8}
    assert p.integer_vars == ["w"]
    assert dict(p.var_bounds) == {"Z": (0, 4), "w": (0, 1), "X": (None, None)}
    assert norm(p.constraints) == {("<=", (("X", 3), ("Y", 1)), 8), ("<=", (("Y", 1), ("Z", 2)), 10),
                                   ("<=", (("X", -2), ("Z", 1), ("w", -1)), 1)}
    with open(os.path.join(DATA, "advanced-problem.mps")) as f:
        q = X.read_mps(f, None, read_case="preserve")             # first RHS vector by default
    assert norm(q.constraints) == {("<=", (("X", 3), ("Y", 1)), 10), ("<=", (("Y", 1), ("Z", 2)), 18),
                                   (">=", (("X", 2), ("Z", -1), ("w", 1)), 6)}


@pytest.mark.parametrize("mode,want", [("upcase", {"W", "X", "Y", "Z"}), ("downcase", {"w", "x", "y", "z"}),
                                       ("invert", {"W", "x", "y", "z"})])
def test_read_mps_read_case(mode, want):
    """t/external-formats.lisp:254-287"""
    with open(os.path.join(DATA, "advanced-problem.mps")) as f:
        assert set(X.read_mps(f, None, read_case=mode, rhs_id="rhs1").vars) == want


def test_read_mps_ranges_single_variable_rows_and_errors():
    text = """NAME          r
ROWS
 N  cost
 L  cap
 G  floor
 E  fix
 L  onlyx
COLUMNS
    a         cost      2.5             cap       1
    a         floor     1               fix       1
    a         onlyx     2
    b         cost      1e1             cap       1
    b         floor     1               fix       -1
RHS
    r         cap       10              floor     2
    r         fix       0               onlyx     6
RANGES
    r         cap       4               floor     3
BOUNDS
 MI bnd       b
 UI bnd       a         9
ENDATA
"""
    p = X.read_mps(text, "min")
    assert dict(p.objective_func) == {"a": Fraction(5, 2), "b": 10}
    assert norm(p.constraints) == {
        ("<=", (("a", 1), ("b", 1)), 10), (">=", (("a", 1), ("b", 1)), 6),      # cap in [10-4, 10]
        (">=", (("a", 1), ("b", 1)), 2), ("<=", (("a", 1), ("b", 1)), 5),       # floor in [2, 2+3]
        ("=", (("a", 1), ("b", -1)), 0)}
    assert dict(p.var_bounds) == {"a": (0, 3), "b": (None, None)}                # 2a <= 6 tightened UI 9
    assert p.integer_vars == ["a"]
    with pytest.raises(conditions.ParsingError):
        X.read_mps(text)                                                         # no problem type
    with pytest.raises(conditions.ParsingError):
        X.read_mps(text.replace(" MI bnd", " QQ bnd"), "min")


def _mps(body):
    return "NAME          bad\nROWS\n N  cost\n L  lim1\n" + body + "ENDATA\n"


@pytest.mark.parametrize("body,needle", [
    ("COLUMNS\n    x         cost               1.0   nosuch             2.0\n", "unknown row"),
    ("COLUMNS\n    x         cost               1.0   lim1               2.x\n", "bad number"),
    ("COLUMNS\n    x         cost               1.0   lim1               1.0\n"
     "RHS\n    rhs       lim1                    \n", "bad number"),
    ("COLUMNS\n    x         cost               1.0   lim1               0.0\n"
     "RHS\n    rhs       lim1               4.0\n", "zero coefficient"),
])
def test_read_mps_malformed_input_raises_parsing_error_not_a_python_error(body, needle):
    """Every malformed record ends in the reference's parsing-error (src/conditions.lisp), with the
    offending line, never in a bare KeyError / ValueError / ZeroDivisionError."""
    with pytest.raises(conditions.ParsingError, match=needle):
        X.read_mps(io.StringIO(_mps(body)), "max")


def test_read_mps_rhs_id_is_compared_as_given():
    """(string= rhs-id current-rhs-id), src/external-formats.lisp:215-218: the set named by
    :rhs-id is used, others are ignored; the comparison does not fold the caller's string."""
    body = ("COLUMNS\n    x         cost               1.0   lim1               1.0\n"
            "    y         cost               1.0   lim1               1.0\n"
            "RHS\n    rhsa      lim1               4.0\n    rhsb      lim1               9.0\n")
    assert X.read_mps(_mps(body), "max", rhs_id="rhsb").constraints[0][2] == 9
    assert X.read_mps(_mps(body), "max").constraints[0][2] == 4          # default: the first set
    assert X.read_mps(_mps(body), "max", rhs_id="RHSB").constraints[0][2] == 0   # no such set


def test_write_standard_format():
    """t/external-formats.lisp:308-337"""
    p = P.make_linear_problem("(max (+ x y))", "(<= (+ (* 2 x) y) 5)")
    out = io.StringIO()
    X.write_standard_format(out, p)
    text = out.getvalue()
    assert text.startswith("Maximize ") and "x" in text and "y" in text and "≤" in text
    assert "<" not in text and "integer" not in text.lower() and "Subject to:" in text
    out = io.StringIO()
    X.write_standard_format(out, p, unicodep=False)
    assert "≤" not in out.getvalue() and "<" in out.getvalue()
    p = P.make_linear_problem("(min (+ x y))", "(<= (+ (* 2 x) y) 5)", "(integer x y)", "(bounds (1 x 7))")
    out = io.StringIO()
    X.write_standard_format(out, p)
    text = out.getvalue()
    assert text.startswith("Minimize ") and "x, y integer" in text and "x ≥ 1" in text and "x ≤ 7" in text


@pytest.mark.gpu
def test_mps_problem_solves_on_the_gpu():
    with open(os.path.join(DATA, "simple-problem.mps")) as f:
        p = X.read_mps(f, "max")
    sol = solver.solve_problem(p)
    assert abs(solver.solution_objective_value(sol) - 92 / 3) <= 1e-8 * 92 / 3
    assert abs(solver.solution_variable(sol, "x") - 8 / 3) < 1e-12
    assert abs(solver.solution_variable(sol, "z") - 3.5) < 1e-12
