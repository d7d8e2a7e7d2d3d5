"""The reference's solver-level tests (t/simplex.lisp, t/solver.lisp, t/integration.lisp) through
the `*solver*` hook with the real backend: DSL -> build-tableau -> C ABI -> sm_100a kernels."""
import pytest

import reference_cases as RC
from linear_programming_b200 import problem as P, simplex, solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", RC.ALL_CASES, ids=lambda c: c.__name__)
def test_reference_case_on_the_gpu(case):
    assert solver.SOLVER is simplex.b200_solver
    case()


def test_pivot_row_is_non_destructive():
    """t/simplex.lisp:135-159"""
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *RC.MAIN)
    t0 = simplex.build_tableau(p, p)
    t1 = simplex.pivot_row(t0, 0, 0)
    assert t0.matrix[0].tolist() == [2, 1, 0, 1, 0, 8] and t0.basis_columns.tolist() == [3, 4]
    assert t1.matrix.tolist() == [[1, .5, 0, .5, 0, 4], [0, 1, 1, 0, 1, 7], [0, -3.5, -3, .5, 0, 4]]
    assert t1.basis_columns.tolist() == [0, 4] and simplex.tableau_objective_value(t1) == 4


def test_backend_keywords():
    p = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *RC.MAIN)
    sol = solver.solve_problem(p, devices=[0], pivot_rule=1)
    assert solver.solution_objective_value(sol) == 28.5
    from linear_programming_b200 import conditions
    with pytest.raises(conditions.SolverError):
        solver.solve_problem(p, max_iterations=1)


def test_random_dsl_problems_agree_with_highs_on_the_gpu():
    """The same 300 random LPs (every constraint and bound kind, two-phase included) through the
    real backend: outcome and objective must match HiGHS."""
    import collections
    import random_problems
    verdicts = collections.Counter(random_problems.check(seed) for seed in range(300))
    assert verdicts["optimal"] > 100 and verdicts["infeasible"] > 10 and verdicts["unbounded"] > 5, verdicts


def test_random_integer_problems_on_the_gpu():
    """Branch and bound with every node's relaxation solved on the B200 (pooled handles)."""
    import random_problems
    for seed in range(40):
        random_problems.check_integer(seed)


def test_ratio_coefficient_two_phase_problems_agree_with_highs_on_the_gpu():
    """600 LPs with k/3, k/7, k/10 data and >= / = rows through the real backend: with the scaled
    feasibility / clean-up tests (the default) verdict and objective match HiGHS on every one."""
    import collections
    import random_problems
    verdicts = collections.Counter(random_problems.check_ratio(seed) for seed in range(600))
    assert verdicts["optimal"] > 300 and verdicts["infeasible"] > 10, verdicts


def test_redundant_equality_row_is_kept_not_an_error():
    """A duplicated `=` row leaves an artificial basic at level zero with nothing to replace it.
    The reference raises a plain `error` there (src/simplex.lisp:432-433, reproduced by
    feas_mode=FEAS_REFERENCE); the default treats the row as redundant and solves the LP."""
    from linear_programming_b200 import _ffi, conditions
    forms = ["(= (+ x y) 4)", "(= (+ x y) 4)", "(<= (+ x (* 2 y)) 6)"]
    p = P.make_linear_problem("(max (+ x y))", *forms)
    sol = solver.solve_problem(p)
    assert solver.solution_objective_value(sol) == 4.0
    with pytest.raises(conditions.SolverError, match="cannot be replaced"):
        solver.solve_problem(p, feas_mode=_ffi.FEAS_REFERENCE)
