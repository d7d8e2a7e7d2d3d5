"""The reference's own solver-level tests, transcribed once and run twice: on the GPU through the
C ABI (tests/test_gpu_solver_hook.py) and on CPU with the oracle standing in for the device
(tests/test_host_frontend.py, which checks the host logic around the hot path only).

Sources: t/simplex.lisp:170-389, t/solver.lisp:20-127, t/integration.lisp:32-124.  The reference
compares exact rationals with `=`; the backend is fp64, so equalities become |a-b| <= 1e-8
relative (BASELINE.json north_star) unless the value is exactly representable and reached exactly.
"""
import math

import pytest

from linear_programming_b200 import conditions, problem as P, simplex, solver

MAIN = ["(<= (+ (* 2 x) y) 8)", "(<= (+ y z) 7)"]


def close(a, b, rel=1e-8):
    return math.isclose(float(a), float(b), rel_tol=rel, abs_tol=1e-9)


def solved(objective, *constraints):
    problem = P.make_linear_problem(objective, *constraints)
    return problem, simplex.n_solve_tableau(simplex.build_tableau(problem, problem))


# ------------------------------------------------------------------ t/simplex.lisp
def case_solve_tableau_basic():
    """t/simplex.lisp:170-194"""
    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    tab0 = simplex.build_tableau(problem, problem)
    tab = simplex.solve_tableau(tab0)
    assert tab is not tab0 and tab0.matrix[-1, -1] == 0.0            # original unchanged
    assert simplex.tableau_objective_value(tab) == 28.5
    assert tab.matrix.tolist() == [[1, 0, -0.5, 0.5, -0.5, 0.5], [0, 1, 1, 0, 1, 7],
                                   [0, 0, 0.5, 0.5, 3.5, 28.5]]
    assert tab.basis_columns.tolist() == [0, 1]


def case_unsolvable_problems():
    """t/simplex.lisp:277-289"""
    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", "(<= (+ (* 2 x) y) 4)",
                                    "(<= (+ y z) 2)", "(>= (+ x z) 5)")
    with pytest.raises(conditions.InfeasibleProblemError):
        simplex.solve_tableau(simplex.build_tableau(problem, problem))
    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", "(<= (+ (* 2 x) y) 4)",
                                    "(<= (+ y (* -1 z)) 4)")
    with pytest.raises(conditions.UnboundedProblemError):
        simplex.solve_tableau(simplex.build_tableau(problem, problem))


def case_two_phase_equality():
    """t/simplex.lisp:196-237"""
    _, tab = solved("(max (+ x (* 4 y) (* 3 z)))", *MAIN, "(= (+ (* 2 x) y z) 8)")
    assert simplex.tableau_objective_value(tab) == 28.5
    assert tab.basis_columns.tolist() == [0, 1, 2]
    assert [simplex.tableau_variable(tab, v) for v in "xyz"] == [0.5, 7.0, 0.0]


def case_two_phase_geq():
    """t/simplex.lisp:239-275"""
    _, tab = solved("(max (+ x (* 4 y) (* 3 z)))", *MAIN, "(>= (+ x z) 1)")
    assert close(simplex.tableau_objective_value(tab), 85 / 3)
    assert tab.basis_columns.tolist() == [1, 2, 0]
    got = [simplex.tableau_variable(tab, v) for v in "xyz"]
    assert all(close(g, w) for g, w in zip(got, (2 / 3, 20 / 3, 1 / 3))), got


def case_tableau_variable():
    """t/simplex.lisp:309-369"""
    obj = "(max (= w (+ x (* 4 y) (* 3 z))))"
    for extra, want in [((), (28.5, 0.5, 7, 0)),
                        (("(bounds (x))",), (28.5, 0.5, 7, 0)),
                        (("(bounds (x 5))",), (28.5, 0.5, 7, 0)),
                        (("(bounds (1 x))",), (28, 1, 6, 1)),
                        (("(bounds (0 y 5))",), (27.5, 1.5, 5, 2))]:
        _, tab = solved(obj, *MAIN, *extra)
        got = [simplex.tableau_variable(tab, v) for v in "wxyz"]
        assert all(close(g, w) for g, w in zip(got, want)), (extra, got)
    with pytest.raises(KeyError):
        simplex.tableau_variable(tab, "foo")
    _, tab = solved("(max (= w (+ (- x) (* 4 y) (* 3 z))))", "(<= (+ (* -2 x) y) 8)",
                    "(<= (+ y z) 7)", "(bounds (x))")
    got = [simplex.tableau_variable(tab, v) for v in "wxyz"]
    assert all(close(g, w) for g, w in zip(got, (28.5, -0.5, 7, 0))), got


def case_tableau_reduced_cost():
    """t/simplex.lisp:372-389"""
    _, tab = solved("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    assert [simplex.tableau_reduced_cost(tab, v) for v in "xyz"] == [0, 0, 0.5]
    with pytest.raises(KeyError):
        simplex.tableau_reduced_cost(tab, "foo")
    _, tab = solved("(max (+ x (* 4 y) (* 3 z)))", *MAIN, "(bounds (z))")
    assert close(simplex.tableau_reduced_cost(tab, "x"), 1)
    assert close(simplex.tableau_reduced_cost(tab, "y"), 0)
    with pytest.raises(ValueError):
        simplex.tableau_reduced_cost(tab, "z")


def case_with_tableau_variables():
    """t/simplex.lisp:391-405"""
    problem, tab = solved("(= w (max (+ x (* 4 y) (* 3 z))))", *MAIN)
    assert simplex.with_tableau_variables("xyzw", tab) == dict(x=0.5, y=7, z=0, w=28.5)
    assert simplex.with_tableau_variables(problem, tab) == dict(w=28.5, x=0.5, y=7, z=0)


# ------------------------------------------------------------------ t/solver.lisp
def case_solve_problem():
    """t/solver.lisp:20-83"""
    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    sol = solver.solve_problem(problem)
    assert solver.solution_problem(sol) is problem
    assert solver.solution_objective_value(sol) == 28.5
    assert [solver.solution_variable(sol, v) for v in "xyz"] == [0.5, 7, 0]
    assert [solver.solution_reduced_cost(sol, v) for v in "xyz"] == [0, 0, 0.5]

    with pytest.raises(conditions.InfeasibleProblemError):
        solver.solve_problem(P.make_linear_problem("(max (+ x y))", "(<= y x)",
                                                   "(>= y (* 1.2 (+ x .9)))", "(integer x y)"))

    rock = ["(<= (+ x y) 5)", "(<= (+ (* -1 x) y) 0)", "(<= (+ (* 6 x) (* 2 y)) 21)",
            "(integer x y)"]
    for objective, want in [("(max (+ (* 240 x) (* 120 y)))", 840),
                            ("(min (+ (* -240 x) (* -120 y)))", -840)]:
        problem = P.make_linear_problem(objective, *rock)
        sol = solver.solve_problem(problem)
        assert solver.solution_problem(sol) is problem
        assert close(solver.solution_objective_value(sol), want)
        assert close(solver.solution_variable(sol, "x"), 3)
        assert close(solver.solution_variable(sol, "y"), 1)
        assert close(solver.solution_reduced_cost(sol, "x"), 0)
        assert close(solver.solution_reduced_cost(sol, "y"), 0)

    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN, "(>= x 1)")
    sol = solver.solve_problem(problem)
    assert close(solver.solution_objective_value(sol), 28)
    assert [round(solver.solution_variable(sol, v), 9) for v in "xyz"] == [1, 6, 1]
    assert [round(solver.solution_reduced_cost(sol, v), 9) for v in "xyz"] == [1, 0, 0]


def case_solution_variable():
    """t/solver.lisp:85-103"""
    problem = P.make_linear_problem("(max (= w (+ x (* 4 y) (* 3 z))))", *MAIN)
    sol = solver.solve_problem(problem)
    assert [solver.solution_variable(sol, v) for v in "wxyz"] == [28.5, 0.5, 7, 0]
    with pytest.raises(KeyError):
        solver.solution_variable(sol, "v")
    with pytest.raises(KeyError):
        solver.solution_reduced_cost(sol, "w")
    assert [solver.solution_reduced_cost(sol, v) for v in "xyz"] == [0, 0, 0.5]
    with pytest.raises(KeyError):
        solver.solution_reduced_cost(sol, "v")


def case_with_solved_problem():
    """t/solver.lisp:105-127"""
    s = solver.with_solved_problem("(max (= w (+ x (* 4 y) (* 3 z))))", *MAIN)
    assert (s.w, s.x, s.y, s.z) == (28.5, 0.5, 7, 0)
    assert [s.reduced_cost(v) for v in "xyz"] == [0, 0, 0.5]
    problem = P.make_linear_problem("(max (= w (+ x (* 4 y) (* 3 z))))", *MAIN)
    s = solver.with_solution_variables(["w", "x", "z"], solver.solve_problem(problem))
    assert (s.w, s.x, s.z) == (28.5, 0.5, 0) and "y" not in s
    assert s.reduced_cost("x") == 0 and s.reduced_cost("z") == 0.5


def case_fp_tolerance_keyword_reaches_the_backend():
    """src/solver.lisp:53-56 passes keyword args through; src/simplex.lisp:511 reads :fp-tolerance"""
    problem = P.make_linear_problem("(max (+ x (* 4 y) (* 3 z)))", *MAIN)
    sol = solver.solve_problem(problem, fp_tolerance=64)
    assert sol.fp_tolerance_factor == 64 and solver.solution_objective_value(sol) == 28.5


# ------------------------------------------------------------------ t/integration.lisp
def case_integration_basic_problem():
    """t/integration.lisp:32-58"""
    s = solver.with_solved_problem(
        "(= revenue (max (* 3 widgets)))",
        "(<= (+ (* 4 widgets) (* -7 d1) (* -6 d2) (* -8 d3)) 0)",
        "(<= (+ (* 3 widgets) (* -5 d1) (* -9 d2) (* -4 d3)) 0)",
        "(<= (+ (* 8 d1) (* 5 d2) (* 3 d3)) 100)",
        "(<= (+ (* 6 d1) (* 9 d2) (* 8 d3)) 200)")
    assert 136.08 <= s.revenue <= 136.11 and close(s.revenue, 160200 / 1177)
    assert 45.36 <= s.widgets <= 45.37
    assert 2.37 <= s.d1 <= 2.38 and 6.96 <= s.d2 <= 6.97 and 15.37 <= s.d3 <= 15.38
    for v in ("widgets", "d1", "d2", "d3"):
        assert abs(s.reduced_cost(v)) < 1e-12


def case_integration_excessive_constraints():
    """t/integration.lisp:63-69"""
    s = solver.with_solved_problem("(min a)", "(<= 0 (+ 148 (* 49 a)) (* 255 a))",
                                   "(<= 0 (+ 135 (* 49 a)) (* 255 a))",
                                   "(<= 0 (+ 134 (* 49 a)) (* 255 a))", "(<= 0 a 1)")
    assert close(s.a, 74 / 103) and abs(s.reduced_cost("a")) < 1e-12


def case_integration_numerical_issue():
    """t/integration.lisp:74-80 (single-float literals, widened exactly to fp64)"""
    s = solver.with_solved_problem("(= z (min (+ b (* 0.6861807 a))))",
                                   "(>= (+ b (* 0.6861807 a)) 0.9372585)",
                                   "(>= (+ b (* 0.7776901 a)) 0.7461006)",
                                   "(>= (+ b (* 0.14247864 a)) 0.38555977)")
    import numpy as np
    assert abs(s.z - float(np.float32(0.9372585))) <= 1e-6
    assert abs(s.z - (s.b + float(np.float32(0.6861807)) * s.a)) <= 1e-6


def case_integration_ilp_bugs():
    """t/integration.lisp:83-107"""
    s = solver.with_solved_problem(
        "(min w)", "(integer x t185 e t184 d t183 c t182 b t181 a t180 w)", "(bounds (1 x 1))",
        "(= (+ (* -1 x) (* 1 t185)) 0)", "(= (+ (* -1 e) (* 1 t184)) 0)",
        "(= (+ (* -1 d) (* 1 t183)) 0)", "(= (+ (* -1 c) (* 1 t182)) 0)",
        "(= (+ (* -1 b) (* 1 t181)) 0)", "(= (+ (* -1 a) (* 10 t180)) 0)",
        "(<= (+ (* -1 e) (* 1 t185)) 0)", "(<= (+ (* -1 d) (* 1 t184)) 0)",
        "(<= (+ (* -1 c) (* 1 t183)) 0)", "(<= (+ (* -1 b) (* 1 t182)) 0)",
        "(<= (+ (* -1 a) (* 7 t182) (* 7 t183) (* 7 t184) (* 7 t185)) 0)",
        """(<= (+ (* -1 w) (* 171 t1) (* 114 t3) (* 189 t10) (* 121 t15) (* 156 t18)
              (* 185 t52) (* 111 t54) (* 141 t63) (* 156 t72) (* 185 t106) (* 111 t108)
              (* 141 t117) (* 156 t126) (* 185 t160) (* 111 t162) (* 141 t171)
              (* 10 t180) (* 1 t181)) 0)""")
    assert close(s.w, 31)
    s = solver.with_solved_problem("(min (+ x y z))", "(integer x y z)",
                                   "(>= (+ x y (* 9 z)) 30/16)",
                                   "(>= (+ (* 3/2 x) (* 78/64 y) z) 32/11)")
    assert (round(s.x, 9), round(s.y, 9), round(s.z, 9)) == (2, 0, 0)


def case_integration_variable_bounds_bug():
    """t/integration.lisp:111-124"""
    s = solver.with_solved_problem("(min (= w (+ x y)))", "(>= x 1.0)", "(>= y 1.0)",
                                   "(>= (+ x (* 2.0 y)) 2.0)")
    assert s.x == 1.0 and s.y == 1.0
    s = solver.with_solved_problem("(min (= w (+ x y)))", "(>= x 1.0)", "(>= y 1.0)")
    assert s.x == 1.0 and s.y == 1.0


ALL_CASES = [v for k, v in sorted(globals().items()) if k.startswith("case_")]
