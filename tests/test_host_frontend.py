"""CPU tests of the host side above the C ABI: DSL parsing, build-tableau (against the reference's
golden initial tableaus), accessors, the `*solver*` hook and branch and bound.

The hot path itself (n-solve-tableau) only exists on the GPU.  To exercise the HOST logic around
it without a GPU, these tests replace `simplex.n_solve_tableau` with the oracle -- the oracle is
the checker's stand-in device here, never a product fallback (the unpatched function raises when
libb200lp.so or a GPU is missing, see test_no_cpu_fallback)."""
import numpy as np
import pytest

import reference_cases as RC
from golden import reference_goldens as G
from linear_programming_b200 import _ffi, conditions, problem as P, sexp, simplex, solver
from oracle import oracle


def f64(rows):
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float64)


def oracle_n_solve_tableau(tableau, **backend):
    if isinstance(tableau, (list, tuple)):
        art, main = tableau
        st, _ = oracle.solve_two_phase(art.matrix, art.basis_columns, main.matrix,
                                       main.basis_columns, main.instance_problem.type == "max",
                                       tol=float(main.fp_tolerance_factor),
                                       feas_mode=backend.get("feas_mode", oracle.FEAS_SCALED))
        conditions.raise_for_status(st)
        return main
    st, _, _ = oracle.solve(tableau.matrix, tableau.basis_columns,
                            tableau.instance_problem.type == "max",
                            tol=float(tableau.fp_tolerance_factor))
    conditions.raise_for_status(st)
    return tableau


@pytest.fixture
def oracle_device(monkeypatch):
    monkeypatch.setattr(simplex, "n_solve_tableau", oracle_n_solve_tableau)


MAIN = RC.MAIN


# ------------------------------------------------------------------------------ reader / parser
def test_sexp_reader_numbers_follow_the_lisp_reader():
    assert sexp.read("(max (+ x (* 4 y) 30/16 1.5d0 .5 -7))") == \
        ["max", ["+", "x", ["*", 4, "y"], sexp.Fraction(15, 8), 1.5, 0.5, -7]]
    assert sexp.read("0.6861807") == float(np.float32(0.6861807))       # single-float literal
    assert sexp.read("(bounds (x) (1 Y nil))") == ["bounds", ["x"], [1, "y", None]]


def test_make_linear_problem_errors():
    """t/problem.lisp:16-32"""
    for objective, cons in [("(avg (+ x (* 4 y) (* 8 z)))", ["(<= (+ x y) 8)"]),
                            ("(min (+ x (* 4 y) (* 8 z)))", ["(& (+ x y) 8)"]),
                            ("(min (+ x (* 4 y) (* 8 z)))", ["(<= (+ x y) 8)", "(foobar x)"])]:
        with pytest.raises(conditions.ParsingError):
            P.make_linear_problem(objective, *cons)
    with pytest.raises(conditions.NonlinearError):
        P.make_linear_problem("(max (* x y))", "(<= x 1)")
    with pytest.raises(conditions.InvalidBoundsError):
        P.make_linear_problem("(max x)", "(<= (+ x y) 4)", "(bounds (3 x 1))")


def test_make_linear_problem_structure():
    """t/problem.lisp:34-75: constraints are normalised to (op alist rhs>=0); single-variable
    constraints become bounds."""
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 8 z)))", "(<= (+ (* 2 x) y) 8)",
                              "(<= (+ y z) 7)", "(>= (+ x z) 1)", "(<= x y)")
    assert p.type == "max" and set(p.vars) == {"x", "y", "z"}
    assert dict(p.objective_func) == dict(x=1, y=4, z=8) and not p.integer_vars and not p.var_bounds
    norm = {(op, tuple(sorted(t)), rhs) for op, t, rhs in p.constraints}
    assert norm == {("<=", (("x", 2), ("y", 1)), 8), ("<=", (("y", 1), ("z", 1)), 7),
                    (">=", (("x", 1), ("z", 1)), 1), ("<=", (("x", 1), ("y", -1)), 0)}
    p = P.make_linear_problem("(= w (min (+ x y)))", "(>= x 1)", "(<= 0 y 5)", "(binary b)",
                              "(<= (+ x y b) 9)")
    assert p.objective_var == "w" and p.type == "min"
    assert dict(p.var_bounds) == dict(x=(1, None), y=(0, 5), b=(0, 1)) and p.integer_vars == ["b"]
    assert len(p.constraints) == 1


# ------------------------------------------------------------------------------ build-tableau
def test_build_tableau_rejects_unknown_operator():
    """t/simplex.lisp:46-57"""
    p = P.Problem(type="max", vars=("x", "y"), objective_var="z",
                  objective_func=[("x", 1), ("y", 2)],
                  constraints=[("<=", [("x", 5), ("y", 1)], 10), ("/=", [("x", 1), ("y", 1)], 5)])
    with pytest.raises(conditions.ParsingError):
        simplex.build_tableau(p, p)


def test_build_tableau_basic_golden():
    """t/simplex.lisp:59-72"""
    g = G.BASIC_INITIAL
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    t = simplex.build_tableau(p, p)
    assert isinstance(t, simplex.Tableau) and t.problem is p and t.instance_problem is p
    assert np.array_equal(t.matrix, f64(g["matrix"])) and t.matrix.dtype == np.float64
    assert t.basis_columns.tolist() == g["basis"] and t.basis_columns.dtype == np.int32
    assert (t.var_count, t.constraint_count) == (g["var_count"], g["constraint_count"])
    assert simplex.tableau_objective_value(t) == 0


@pytest.mark.parametrize("extra,g", [("(= (+ (* 2 x) y z) 8)", G.EQ_BUILD),
                                     ("(>= (+ x z) 1)", G.GEQ_BUILD)], ids=["eq", "geq"])
def test_build_tableau_two_phase_goldens(extra, g):
    """t/simplex.lisp:74-133"""
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN, extra)
    art, main = simplex.build_tableau(p, p)
    assert np.array_equal(art.matrix, f64(g["art_matrix"]))
    assert art.basis_columns.tolist() == g["art_basis"] and art.var_count == g["art_var_count"]
    assert simplex.tableau_objective_value(art) == g["art_objective"]
    assert art.instance_problem.type == "min" and art.problem is p
    assert np.array_equal(main.matrix, f64(g["main_matrix"]))
    assert main.basis_columns.tolist() == g["main_basis"] and main.var_count == g["main_var_count"]
    assert main.constraint_count == art.constraint_count == g["constraint_count"]


def test_build_tableau_bounds_and_negative_rhs():
    """src/simplex.lisp:189-212 (bound kinds), :243-252 (row negation)."""
    p = P.make_linear_problem("(max (+ x y u f))", "(<= (+ x y u f) 10)", "(<= (+ x (* -1 y)) -2)",
                              "(bounds (1 x) (0 y 5) (u 7) (f))", "(<= u 7)")
    art, main = simplex.build_tableau(p, p)
    vm = main.var_mapping
    assert vm["x"] == ("positive", 0, 1) and vm["y"] == ("positive", 1, 0)
    assert vm["u"] == ("negative", 2, 7) and vm["f"] == ("signed", 3)
    # y's upper bound became a row pushed IN FRONT of the constraints
    assert main.matrix[0].tolist() == [0, 1, 0, 0, 0, 1, 0, 0, 5]
    # x + y + u + f <= 10 with x -> x'+1, u -> 7-u':  x' + y - u' + f+ - f- <= 2
    assert main.matrix[1].tolist() == [1, 1, -1, 1, -1, 0, 1, 0, 2]
    # x - y <= -2 -> x' - y <= -3 -> negated: -x' + y >= 3 (surplus -1, artificial)
    assert main.matrix[2].tolist() == [-1, 1, 0, 0, 0, 0, 0, -1, 3]
    assert main.basis_columns.tolist() == [5, 6, 9]
    # objective row: -coef for positive/signed+, +coef for negative; offsets in the corner
    assert main.matrix[3].tolist() == [-1, -1, 1, -1, 1, 0, 0, 0, 8]
    assert art.matrix[3].tolist() == [-1, 1, 0, 0, 0, 0, 0, -1, 0, 3]


def test_build_tableau_without_constraints():
    """src/simplex.lisp:153-186"""
    p = P.make_linear_problem("(max (+ x (* -2 y)))", "(bounds (1 x 4) (2 y 9))")
    t = simplex.build_tableau(p, p)
    assert np.array_equal(t.matrix, f64([[1, 0, 0], [0, 1, 0], [0, 0, 0]]))
    assert simplex.tableau_variable(t, "x") == 4 and simplex.tableau_variable(t, "y") == 2
    assert simplex.tableau_objective_value(t) == 0     # the corner holds 4 - 4 = 0
    p = P.make_linear_problem("(max x)", "(>= x 1)")
    with pytest.raises(conditions.UnboundedProblemError):
        simplex.build_tableau(p, p)


def test_copy_tableau():
    """t/simplex.lisp:293-307"""
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    t1 = simplex.build_tableau(p, p)
    t2 = simplex.copy_tableau(t1)
    assert t1 is not t2 and t1.problem is t2.problem
    assert t1.matrix is not t2.matrix and np.array_equal(t1.matrix, t2.matrix)
    assert t1.basis_columns is not t2.basis_columns
    assert np.array_equal(t1.basis_columns, t2.basis_columns)
    assert (t1.var_count, t1.constraint_count) == (t2.var_count, t2.constraint_count)


# --------------------------------------------- the reference's solver-level tests, host logic only
@pytest.mark.parametrize("case", RC.ALL_CASES, ids=lambda c: c.__name__)
def test_reference_case_with_oracle_standing_in_for_the_device(case, oracle_device):
    case()


def test_solver_hook_is_rebindable():
    """src/solver.lisp:39-56: *solver* is any function (problem &key ...) -> solution."""
    seen = {}

    class Sol:
        pass

    def fake(problem, **kw):
        seen.update(kw, problem=problem)
        return Sol()

    solver.solution_objective_value.register(Sol, lambda s: 42)
    p = P.make_linear_problem("(max x)", "(<= (+ x y) 3)")
    with solver.using_solver(fake):
        sol = solver.solve_problem(p, fp_tolerance=7, custom=1)
        assert solver.solution_objective_value(sol) == 42
    assert seen == dict(problem=p, fp_tolerance=7, custom=1)
    assert solver.SOLVER is simplex.b200_solver          # binding restored; default = the drop-in
    with pytest.raises(TypeError):
        solver.solution_variable(object(), "x")


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the host."""
    if _ffi.device_count() > 0:
        pytest.skip("a GPU is present")
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    with pytest.raises(_ffi.B200DeviceError):
        solver.solve_problem(p)


def test_status_codes_map_onto_the_reference_conditions():
    """include/b200lp.h status codes -> src/conditions.lisp:43-77"""
    conditions.raise_for_status(0)
    for code, exc in [(1, conditions.UnboundedProblemError), (2, conditions.InfeasibleProblemError),
                      (3, conditions.SolverError), (4, conditions.SolverError),
                      (5, conditions.SolverError)]:
        with pytest.raises(exc):
            conditions.raise_for_status(code)
    assert issubclass(conditions.InfeasibleIntegerConstraintsError, conditions.InfeasibleProblemError)
    assert issubclass(conditions.UnboundedProblemError, conditions.SolverError)


def test_random_dsl_problems_agree_with_highs(oracle_device):
    """parse -> build-tableau (bounds, negated rows, artificials) -> solve -> accessors, against an
    independent solver, on 300 random small LPs; the oracle stands in for the device."""
    import collections
    import random_problems
    verdicts = collections.Counter(random_problems.check(seed) for seed in range(300))
    assert verdicts["optimal"] > 100 and verdicts["infeasible"] > 10 and verdicts["unbounded"] > 5, verdicts
    assert verdicts["stuck"] <= 3, verdicts


def test_random_integer_problems_agree_with_a_milp_solver(oracle_device):
    """Branch and bound (src/simplex.lisp:466-542 as b200_solver) on 80 random pure-integer LPs."""
    import random_problems
    for seed in range(80):
        random_problems.check_integer(seed)


def test_ratio_coefficient_two_phase_problems_agree_with_highs(oracle_device):
    """LPs with k/3, k/7, k/10 data and >= / = rows (phase 1 on every one): the reference solves
    them exactly; an fp64 backend ends phase 1 at a residue of a few 1e-13, which the reference's
    ABSOLUTE 1024*eps feasibility test and exact `/= 0` clean-up tests then misjudge.  With the
    backend's scaled tests (B200LP_FEAS_SCALED, the default) every verdict and objective agrees
    with HiGHS; with the literal ones (B200LP_FEAS_REFERENCE) some feasible LPs are rejected --
    kept visible here so the deviation stays a documented choice."""
    import collections
    import random_problems
    verdicts = collections.Counter(random_problems.check_ratio(seed) for seed in range(600))
    assert verdicts["optimal"] > 300 and verdicts["infeasible"] > 10, verdicts
    wrong = 0
    for seed in range(600):
        objective, forms, names, ref = random_problems.generate_ratio(seed)
        if ref.status != 0:
            continue
        try:
            solver.solve_problem(P.make_linear_problem(objective, *forms), feas_mode=oracle.FEAS_REFERENCE)
        except conditions.SolverError:
            wrong += 1
    assert wrong >= 1, "the literal absolute tests no longer misjudge any fp64 residue?"


def test_integrality_test_tolerates_fp64_noise_but_not_fractions():
    """simplex._is_integral: 1e-9 * max(1, |v|), never tighter than factor * eps."""
    assert simplex._is_integral(3.0000000000004, 1024) and simplex._is_integral(1e6 + 2e-5, 1024)
    assert not simplex._is_integral(3.00001, 1024) and not simplex._is_integral(0.5, 1024)
    assert simplex._is_integral(7.0, 1024) and simplex._is_integral(-2.0 - 1e-12, 1024)


def test_dense_block_equals_the_rows_of_the_full_generator():
    """synthetic.dense_block (what a rank of a sharded bench uploads) skips the other ranks' rows by
    advancing PCG64; the block must be bit-identical to the same rows of dense_tableau."""
    import numpy as np
    from linear_programming_b200 import synthetic
    full, basis = synthetic.dense_tableau(37, 53)
    for lo, hi in [(0, 37), (5, 20), (30, 37), (0, 1)]:
        blk, bb = synthetic.dense_block(37, 53, lo, hi)
        assert np.array_equal(blk[:-1], full[lo:hi]) and np.array_equal(blk[-1], full[-1])
        assert np.array_equal(bb, basis[lo:hi])
