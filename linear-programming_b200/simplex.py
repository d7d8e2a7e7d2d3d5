"""Host side of the B200 simplex backend: the `tableau` structure, its accessors, `build_tableau`
and the backend function `b200_solver` that `*solver*` is bound to.

Mirrors the reference's src/simplex.lisp everywhere EXCEPT the hot path: `n-solve-tableau`
(:399-461) is not restated here -- `solve_tableau` / `n_solve_tableau` hand the fp64 tableau to
libb200lp.so through the C ABI (include/b200lp.h) and fail loudly when the CUDA library or a GPU
is missing.  There is no CPU fallback.

  tableau struct / copy-tableau / accessors    src/simplex.lisp:44-120
  build-tableau                                src/simplex.lisp:142-328  (writes fp64 directly)
  solve-tableau / n-solve-tableau / pivot-row  src/simplex.lisp:333-461  -> b200lp_solve[_two_phase]
  branch and bound + simplex-solver            src/simplex.lisp:466-542  -> b200_solver
"""
import math
from dataclasses import dataclass, field, replace
from fractions import Fraction

import numpy as np

from . import _ffi
from .conditions import (InfeasibleProblemError, ParsingError, SolverError,
                         UnboundedProblemError, raise_for_status)
from .problem import Problem
from .utils import DOUBLE_FLOAT_EPSILON as CL_DOUBLE_FLOAT_EPSILON


@dataclass
class Tableau:
    """src/simplex.lisp:48-58.  matrix: row-major fp64 (constraint_count+1) x (var_count+1), last
    column = right-hand side, last row = objective row; basis_columns[i] = column basic in row i;
    var_mapping[var] = ('positive'|'negative', column, offset) | ('signed', column)."""
    problem: Problem
    instance_problem: Problem
    matrix: np.ndarray
    basis_columns: np.ndarray
    var_count: int
    constraint_count: int
    var_mapping: dict = field(default_factory=dict)
    fp_tolerance_factor: float = 1024


def copy_tableau(tableau):
    """src/simplex.lisp:61-71: new matrix and basis, everything else shared."""
    return replace(tableau, matrix=tableau.matrix.copy(), basis_columns=tableau.basis_columns.copy())


def tableau_objective_value(tableau):
    """src/simplex.lisp:74-78"""
    return float(tableau.matrix[tableau.constraint_count, tableau.var_count])


def _basic_value(tableau, column):
    """(if-let (idx (position column basis)) matrix[idx, var-count] 0), :93-95"""
    idx = np.flatnonzero(tableau.basis_columns == column)
    return float(tableau.matrix[idx[0], tableau.var_count]) if idx.size else 0


def tableau_variable(tableau, var):
    """src/simplex.lisp:81-107"""
    if var == tableau.instance_problem.objective_var:
        return tableau_objective_value(tableau)
    mapping = tableau.var_mapping.get(var)
    if mapping is None:
        raise KeyError(f"{var} is not a variable in the tableau")
    kind = mapping[0]
    if kind == "positive":
        return mapping[2] + _basic_value(tableau, mapping[1])
    if kind == "negative":
        return mapping[2] + -_basic_value(tableau, mapping[1])
    return _basic_value(tableau, mapping[1]) - _basic_value(tableau, mapping[1] + 1)


def tableau_reduced_cost(tableau, var):
    """src/simplex.lisp:111-120"""
    mapping = tableau.var_mapping.get(var)
    if mapping is None:
        raise KeyError(f"{var} is not a variable in the tableau")
    if mapping[0] != "positive":
        raise ValueError(f"{var} has no lower bound")
    return float(tableau.matrix[tableau.constraint_count, mapping[1]])


def with_tableau_variables(var_list, tableau):
    """src/simplex.lisp:125-139 as a dict: a Problem binds the objective variable and every
    problem variable, a sequence binds just those names."""
    if isinstance(var_list, Problem):
        out = {var_list.objective_var: tableau_objective_value(tableau)}
        out.update({v: tableau_variable(tableau, v) for v in var_list.vars})
        return out
    return {v: tableau_variable(tableau, v) for v in var_list}


# ------------------------------------------------------------------------------ build-tableau
def _find_bound(problem, var):
    for name, bound in problem.var_bounds:
        if name == var:
            return bound
    return None


def _f64(x):
    """What the Lisp shim's (coerce x 'double-float) yields for an integer / ratio / float."""
    if isinstance(x, Fraction):
        return x.numerator / x.denominator      # correctly rounded, like CL's coerce of a ratio
    return float(x)


def build_tableau(problem, instance_problem, fp_tolerance_factor=1024):
    """src/simplex.lisp:142-328.  Returns a Tableau, or [art_tableau, main_tableau] when the slack
    basis is not feasible.  Cell values are computed in the input's own numeric type (exact for
    integers / ratios, as in the reference) and written straight into fp64 storage -- the boxed
    (simple-array real 2) of the reference is never materialised."""
    constraints = list(instance_problem.constraints)
    nvars = len(problem.vars)
    mappings = {}
    if not constraints:                                             # :153-186
        matrix = np.zeros((nvars + 1, nvars + 1))
        basis = np.arange(nvars, dtype=np.int32)
        objective = 0
        is_max = problem.type == "max"
        coefs = dict(problem.objective_func)
        for i, var in enumerate(problem.vars):
            coef = coefs.get(var, 0)
            lb, ub = _find_bound(problem, var) or (None, None)
            matrix[i, i] = 1.0
            pick = ub if (0 <= coef) == is_max else lb
            if pick is None:
                raise UnboundedProblemError()
            mappings[var] = ("positive", i, pick)
            objective = objective + coef * pick
        matrix[nvars, nvars] = _f64(objective)
        return Tableau(problem, problem, matrix, basis, nvars, nvars, mappings, fp_tolerance_factor)

    column = 0
    for var in problem.vars:                                        # :189-212
        bound = _find_bound(problem, var)
        lb, ub = bound if bound is not None else (None, None)
        if bound is None:
            mappings[var] = ("positive", column, 0)
        elif lb is not None and ub is not None:
            # kept exactly as the reference writes it (:199-203), including `var >= -ub` for a
            # negative upper bound -- see DESIGN.md section 2
            constraints.insert(0, ("<=", [(var, 1)], ub) if 0 <= ub else (">=", [(var, 1)], -ub))
            mappings[var] = ("positive", column, lb)
        elif lb is not None:
            mappings[var] = ("positive", column, lb)
        elif ub is not None:
            mappings[var] = ("negative", column, ub)
        else:
            mappings[var] = ("signed", column)
            column += 1
        column += 1
    num_var_cols = column

    m = len(constraints)
    num_slack = sum(1 for c in constraints if c[0] != "=")
    num_cols = num_var_cols + num_slack + 1
    matrix = np.zeros((m + 1, num_cols))
    basis = np.zeros(m, dtype=np.int32)
    rows = []                      # sparse exact rows {col: value}; the art objective row sums them
    art_rows = []
    col_offset = 0
    for row, (op, terms, rhs) in enumerate(constraints):            # :223-265
        cells = {}
        for var, coef in terms:
            mp = mappings[var]
            if mp[0] == "positive":
                cells[mp[1]] = coef
                rhs = rhs - coef * mp[2]
            elif mp[0] == "negative":
                cells[mp[1]] = -coef
                rhs = rhs - coef * mp[2]
            else:
                cells[mp[1]] = coef
                cells[mp[1] + 1] = -coef
        if rhs < 0:                                                 # :243-252
            cells = {c: -v for c, v in cells.items()}
            rhs = -rhs
            op = {"<=": ">=", ">=": "<="}.get(op, op)
        if op == "<=":
            cells[num_var_cols + col_offset] = 1
            basis[row] = num_var_cols + col_offset
            col_offset += 1
        elif op == ">=":
            art_rows.append(row)
            cells[num_var_cols + col_offset] = -1
            basis[row] = num_cols
            col_offset += 1
        elif op == "=":
            art_rows.append(row)
            basis[row] = num_cols
        else:
            raise ParsingError(f"{(op, terms, rhs)!r} is not a valid constraint equation")
        cells[num_cols - 1] = rhs
        rows.append(cells)
        for c, v in cells.items():
            matrix[row, c] = _f64(v)

    obj_rhs = 0
    for var, coef in problem.objective_func:                        # :267-279
        mp = mappings[var]
        if mp[0] == "positive":
            matrix[m, mp[1]] = _f64(-coef)
            obj_rhs = obj_rhs + coef * mp[2]
        elif mp[0] == "negative":
            matrix[m, mp[1]] = _f64(coef)
            obj_rhs = obj_rhs + coef * mp[2]
        else:
            matrix[m, mp[1]] = _f64(-coef)
            matrix[m, mp[1] + 1] = _f64(coef)
    matrix[m, num_cols - 1] = _f64(obj_rhs)

    main = Tableau(problem, instance_problem, matrix, basis, num_cols - 1, m, mappings,
                   fp_tolerance_factor)
    if not art_rows:
        return main

    num_art = len(art_rows)                                         # :288-325
    art = np.zeros((m + 1, num_cols + num_art))
    art[:m, :num_cols - 1] = matrix[:m, :num_cols - 1]
    art[:m, -1] = matrix[:m, -1]
    art_basis = basis.copy()
    # the reference pushes rows, so artificial columns are handed out in reverse row order
    for i, row in enumerate(reversed(art_rows)):
        art_basis[row] = num_cols - 1 + i
        art[row, num_cols - 1 + i] = 1.0
    sums = {}
    for row in art_rows:                                            # ascending row order, :303-316
        for c, v in rows[row].items():
            sums[c] = sums.get(c, 0) + v
    for c, v in sums.items():
        art[m, num_cols + num_art - 1 if c == num_cols - 1 else c] = _f64(v)
    art_problem = Problem(type="min", vars=problem.vars)
    art_tab = Tableau(problem, art_problem, art, art_basis, num_cols - 1 + num_art, m, mappings,
                      fp_tolerance_factor)
    return [art_tab, main]


# ---------------------------------------------------------------- the hot path: on the GPU only
def _opts(tableau, backend):
    return _ffi.make_opts(fp_tolerance=tableau.fp_tolerance_factor,
                          pivot_rule=backend.get("pivot_rule", _ffi.RULE_REFERENCE),
                          max_iters=backend.get("max_iterations", 0),
                          devices=backend.get("devices"),
                          writeback_full=backend.get("writeback_full", True),
                          feas_mode=backend.get("feas_mode", _ffi.FEAS_SCALED))


def n_solve_tableau(tableau, **backend):
    """src/simplex.lisp:399-461 -- destructive solve; the work happens in libb200lp.so.
    A [art, main] list runs the two-phase variant and returns the main tableau."""
    if isinstance(tableau, (list, tuple)):
        art, main = tableau
        is_max = main.instance_problem.type == "max"
        status, _ = _ffi.solve_two_phase(art.matrix, art.basis_columns, main.matrix,
                                         main.basis_columns, is_max, _opts(main, backend))
        raise_for_status(status)
        return main
    is_max = tableau.instance_problem.type == "max"
    status, _, _ = _ffi.solve(tableau.matrix, tableau.basis_columns, is_max, _opts(tableau, backend))
    raise_for_status(status)
    return tableau


def solve_tableau(tableau, **backend):
    """src/simplex.lisp:391-397 -- the original tableau(s) are unchanged."""
    if isinstance(tableau, (list, tuple)):
        return n_solve_tableau([copy_tableau(t) for t in tableau], **backend)
    return n_solve_tableau(copy_tableau(tableau), **backend)


def n_pivot_row(tableau, entering_col, changing_row):
    """src/simplex.lisp:337-359 on the device (b200lp_pivot), destructive."""
    R, C = tableau.matrix.shape
    with _ffi.DeviceTableau(R, C, tableau.instance_problem.type == "max") as dev:
        dev.upload(np.ascontiguousarray(tableau.matrix), tableau.basis_columns)
        dev.pivot(entering_col, changing_row)
        mat, basis = dev.download()
    tableau.matrix[...] = mat
    tableau.basis_columns[...] = basis
    return tableau


def pivot_row(tableau, entering_col, changing_row):
    """src/simplex.lisp:333-335"""
    return n_pivot_row(copy_tableau(tableau), entering_col, changing_row)


# ------------------------------------------------------------------------- branch and bound
INTEGRALITY_TOLERANCE = 1e-9


def _is_integral(value, tol_factor):
    """The reference tests (integerp value) on exact rationals (:479).  An fp64 vertex with an
    integer coordinate carries 1e-12..1e-10 of noise (more for values far from 1), so the question
    is asked with a tolerance of its own, relative to the value: 1e-9 * max(1, |value|) (and never
    tighter than the `fp=` tolerance factor * eps).  Branching on noise would add a row, a
    phase 1 and a GPU solve per spurious node, with nothing bounding the depth."""
    tol = max(INTEGRALITY_TOLERANCE * max(1.0, abs(value)), tol_factor * CL_DOUBLE_FLOAT_EPSILON)
    return abs(value - round(value)) <= tol


def violated_integer_constraint(tableau):
    """src/simplex.lisp:475-480"""
    for var in tableau.problem.integer_vars:
        if not _is_integral(tableau_variable(tableau, var), tableau.fp_tolerance_factor):
            return var
    return None


def gen_entries(tableau, entry):
    """src/simplex.lisp:466-473"""
    var = violated_integer_constraint(tableau)
    val = tableau_variable(tableau, var)             # fractional by more than the tolerance
    return [[("<=", [(var, 1)], math.floor(val))] + entry,
            [(">=", [(var, 1)], math.ceil(val))] + entry]


def build_and_solve(problem, extra_constraints, fp_tolerance_factor=1024, **backend):
    """src/simplex.lisp:483-502: an infeasible node is reported as the value 'infeasible'."""
    try:
        instance = problem if not extra_constraints else replace(
            problem, constraints=list(extra_constraints) + list(problem.constraints))
        return solve_tableau(build_tableau(problem, instance, fp_tolerance_factor), **backend)
    except InfeasibleProblemError:
        return "infeasible"


def b200_solver(problem, **kwargs):
    """The `*solver*` backend function (contract: src/solver.lisp:39-56; control flow of
    simplex-solver, src/simplex.lisp:506-542, with every LP relaxation solved on the B200).

    Keywords: fp_tolerance (the reference's :fp-tolerance, default 1024) plus the backend's own
    devices=[...], pivot_rule, max_iterations (allowed by solve-problem's &allow-other-keys)."""
    tol = kwargs.pop("fp_tolerance", 1024)
    backend = {k: kwargs[k] for k in ("devices", "pivot_rule", "max_iterations", "feas_mode")
               if k in kwargs}
    better = (lambda a, b: a < b) if problem.type == "max" else (lambda a, b: a > b)
    best, solution = None, None
    stack = [[]]
    while stack:
        entry = stack.pop(0)
        tab = build_and_solve(problem, entry, tol, **backend)
        if isinstance(tab, str):
            continue                                    # infeasible leaf
        violated = violated_integer_constraint(tab)
        value = tableau_objective_value(tab)
        if violated is not None and best is not None and not better(best, value):
            continue                                    # cannot contain the optimum
        if violated is not None:
            stack = gen_entries(tab, entry) + stack
        elif best is None or better(best, value):
            best, solution = value, tab
    if solution is None:
        raise InfeasibleProblemError()
    return solution


__all__ = ["Tableau", "copy_tableau", "tableau_objective_value", "tableau_variable",
           "tableau_reduced_cost", "with_tableau_variables", "build_tableau", "solve_tableau",
           "n_solve_tableau", "pivot_row", "n_pivot_row", "violated_integer_constraint",
           "gen_entries", "build_and_solve", "b200_solver", "SolverError"]
