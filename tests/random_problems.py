"""Random small LPs in the DSL (mixed <=, >=, = rows, every kind of variable bound, max and min),
with scipy's HiGHS as an independent judge of the outcome.  Shared by the CPU test (oracle standing
in for the device) and the GPU test (real backend)."""
import numpy as np
from scipy.optimize import linprog

from linear_programming_b200 import conditions, problem as P, solver


def generate(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 7))
    m = int(rng.integers(1, 7))
    names = [f"v{i}" for i in range(n)]
    sense = "max" if rng.random() < 0.5 else "min"
    c = rng.integers(-5, 6, size=n)
    forms, A_ub, b_ub, A_eq, b_eq = [], [], [], [], []
    for _ in range(m):
        a = rng.integers(-4, 6, size=n) * (rng.random(n) < 0.8)
        if np.count_nonzero(a) < 2:                       # single-variable rows become bounds: keep them apart
            a[:2] = (1, 1)
        rhs = int(rng.integers(-6, 25))
        op = ("<=", ">=", "=")[int(rng.choice(3, p=[0.6, 0.25, 0.15]))]
        lhs = "(+ " + " ".join(f"(* {int(k)} {v})" for k, v in zip(a, names) if k != 0) + ")"
        forms.append(f"({op} {lhs} {rhs})")
        if op == "<=":
            A_ub.append(a); b_ub.append(rhs)
        elif op == ">=":
            A_ub.append(-a); b_ub.append(-rhs)
        else:
            A_eq.append(a); b_eq.append(rhs)
    bounds, entries = [], []
    for v in names:
        kind = int(rng.integers(0, 6))
        lo, hi = 0, None                                   # implicit x >= 0
        if kind == 1:
            lo = int(rng.integers(-3, 4)); entries.append(f"({lo} {v})")
        elif kind == 2:
            lo, hi = None, int(rng.integers(0, 9)); entries.append(f"({v} {hi})")
        elif kind == 3:
            # hi >= 0 only: for a NEGATIVE upper bound next to a lower bound the reference's
            # build-tableau emits `var >= -ub` (src/simplex.lisp:199-203), a quirk the backend
            # reproduces faithfully and HiGHS naturally does not
            lo = int(rng.integers(-3, 3)); hi = max(0, lo + int(rng.integers(0, 8))); entries.append(f"({lo} {v} {hi})")
        elif kind == 4:
            lo, hi = None, None; entries.append(f"({v})")
        bounds.append((lo, hi))
    if entries:
        forms.append("(bounds " + " ".join(entries) + ")")
    objective = f"({sense} (+ " + " ".join(f"(* {int(k)} {v})" for k, v in zip(c, names)) + "))"
    sign = -1.0 if sense == "max" else 1.0
    kw = dict(A_ub=np.array(A_ub) if A_ub else None, b_ub=b_ub or None,
              A_eq=np.array(A_eq) if A_eq else None, b_eq=b_eq or None, bounds=bounds, method="highs")
    ref = linprog(sign * c, **kw)
    # HiGHS' presolve reports some unbounded problems as "infeasible" (dual infeasibility);
    # a zero-objective solve settles whether a feasible point exists
    ref.feasible = ref.status == 0 or linprog(np.zeros(n), **kw).status == 0
    return objective, forms, names, sign, ref


def check(seed):
    """Returns a short verdict string; raises AssertionError on a disagreement with HiGHS."""
    objective, forms, names, sign, ref = generate(seed)
    problem = P.make_linear_problem(objective, *forms)
    try:
        sol = solver.solve_problem(problem)
    except conditions.InfeasibleProblemError:
        assert not ref.feasible, (seed, "backend: infeasible", ref.status, ref.message)
        return "infeasible"
    except conditions.UnboundedProblemError:
        assert ref.feasible and ref.status != 0, (seed, "backend: unbounded", ref.status, ref.message)
        return "unbounded"
    except conditions.SolverError as exc:                  # the reference's plain `error` on a stuck artificial
        if "Artificial" in str(exc):
            return "stuck"
        raise
    assert ref.status == 0, (seed, "backend: optimal", ref.status, ref.message)
    want = sign * ref.fun
    got = solver.solution_objective_value(sol)
    assert abs(got - want) <= 1e-7 * max(1.0, abs(want)), (seed, got, want)
    x = np.array([solver.solution_variable(sol, v) if v in problem.vars else 0.0 for v in names])
    # the reported point must itself be feasible for HiGHS's statement of the problem
    coefs = dict(problem.objective_func)
    assert abs(sum(coefs.get(v, 0) * xi for v, xi in zip(names, x)) - got) <= 1e-7 * max(1.0, abs(got))
    return "optimal"


def generate_ratio(seed):
    """LPs whose data are RATIOS (k/3, k/7, k/10) with mostly >= and = rows: exact in the
    reference, rounded on entry to an fp64 backend -- the phase-1 objective then ends at a residue
    of a few 1e-13 instead of 0, which an absolute 1024*eps test calls infeasible."""
    from fractions import Fraction
    rng = np.random.default_rng(50_000 + seed)
    n = int(rng.integers(2, 7))
    m = int(rng.integers(2, 6))
    names = [f"r{i}" for i in range(n)]
    dens = (1, 3, 7, 10)

    def frac(lo, hi):
        return Fraction(int(rng.integers(lo, hi)), int(dens[int(rng.integers(0, 4))]))

    c = [frac(1, 20) for _ in range(n)]                    # min with positive costs: bounded below
    forms, A_ub, b_ub, A_eq, b_eq = [], [], [], [], []
    x0 = [frac(0, 30) for _ in range(n)]                   # most rows are built around a feasible point
    for _ in range(m):
        a = [frac(0, 12) if rng.random() < 0.8 else Fraction(0) for _ in range(n)]
        if sum(1 for v in a if v) < 2:
            a[0], a[1] = frac(1, 12), frac(1, 12)
        at_x0 = sum(ai * xi for ai, xi in zip(a, x0))
        op = (">=", "=", "<=")[int(rng.choice(3, p=[0.5, 0.3, 0.2]))]
        rhs = at_x0 if op == "=" else at_x0 - frac(0, 9) if op == ">=" else at_x0 + frac(0, 9)
        if rng.random() < 0.1:
            rhs = rhs + frac(5, 40)                         # sometimes off the feasible point
        rhs = max(rhs, Fraction(0))
        lhs = "(+ " + " ".join(f"(* {k} {v})" for k, v in zip(a, names) if k) + ")"
        forms.append(f"({op} {lhs} {rhs})")
        af = [float(v) for v in a]
        if op == "<=":
            A_ub.append(af); b_ub.append(float(rhs))
        elif op == ">=":
            A_ub.append([-v for v in af]); b_ub.append(-float(rhs))
        else:
            A_eq.append(af); b_eq.append(float(rhs))
    objective = "(min (+ " + " ".join(f"(* {k} {v})" for k, v in zip(c, names)) + "))"
    kw = dict(A_ub=np.array(A_ub) if A_ub else None, b_ub=b_ub or None,
              A_eq=np.array(A_eq) if A_eq else None, b_eq=b_eq or None,
              bounds=[(0, None)] * n, method="highs")
    ref = linprog([float(v) for v in c], **kw)
    return objective, forms, names, ref


def check_ratio(seed):
    objective, forms, names, ref = generate_ratio(seed)
    problem = P.make_linear_problem(objective, *forms)
    try:
        sol = solver.solve_problem(problem)
    except conditions.InfeasibleProblemError:
        assert ref.status == 2, (seed, "backend: infeasible", ref.status, ref.message)
        return "infeasible"
    assert ref.status == 0, (seed, "backend: optimal", ref.status, ref.message)
    got = solver.solution_objective_value(sol)
    assert abs(got - ref.fun) <= 1e-7 * max(1.0, abs(ref.fun)), (seed, got, ref.fun)
    return "optimal"


def check_integer(seed):
    """A random small pure-integer LP through the hook's branch and bound, against scipy's MILP."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    rng = np.random.default_rng(10_000 + seed)
    n = int(rng.integers(2, 5))
    m = int(rng.integers(2, 5))
    names = [f"k{i}" for i in range(n)]
    sense = "max" if rng.random() < 0.7 else "min"
    c = rng.integers(1, 9, size=n) * (1 if sense == "max" else -1) * rng.choice([1, 1, 1, -1], size=n)
    A = rng.integers(0, 7, size=(m, n))
    A[:, 0] = np.maximum(A[:, 0], 1)
    A[0] = np.maximum(A[0], 1)                              # every variable is capped by row 0
    b = rng.integers(5, 40, size=m)
    forms = ["(<= (+ " + " ".join(f"(* {int(k)} {v})" for k, v in zip(row, names) if k) + f") {int(r)})"
             for row, r in zip(A, b)]
    forms.append("(integer " + " ".join(names) + ")")
    objective = f"({sense} (+ " + " ".join(f"(* {int(k)} {v})" for k, v in zip(c, names)) + "))"
    sign = -1.0 if sense == "max" else 1.0
    ref = milp(sign * c, constraints=LinearConstraint(A, ub=b), integrality=np.ones(n),
               bounds=Bounds(0, np.inf))
    problem = P.make_linear_problem(objective, *forms)
    sol = solver.solve_problem(problem)
    assert ref.status == 0, (seed, ref.message)
    got, want = solver.solution_objective_value(sol), sign * ref.fun
    assert abs(got - want) <= 1e-7 * max(1.0, abs(want)), (seed, got, want)
    x = [solver.solution_variable(sol, v) for v in names]
    assert all(abs(xi - round(xi)) < 1e-9 for xi in x), (seed, x)
    assert (A @ np.round(x) <= b + 1e-9).all(), (seed, x)
    return got
