"""The high-level solver interface (reference: src/solver.lisp) with the B200 backend installed.

`SOLVER` plays the role of the special variable `*solver*` (src/solver.lisp:39-49): a function
(problem, **keyword_args) -> solution object answering the four solution_* generics.  In the
reference the default is the pure-Lisp `simplex-solver`; in this package the default IS the
drop-in, `simplex.b200_solver`.  `using_solver` gives the dynamic-binding (`let`) behaviour.
"""
import contextlib
import functools

from . import simplex
from .problem import Problem, parse_linear_problem

SOLVER = simplex.b200_solver


def set_solver(fn):
    """(setf *solver* fn)"""
    global SOLVER
    SOLVER = fn


@contextlib.contextmanager
def using_solver(fn):
    """(let ((*solver* fn)) ...)"""
    global SOLVER
    old, SOLVER = SOLVER, fn
    try:
        yield
    finally:
        SOLVER = old


def solve_problem(problem, **kwargs):
    """src/solver.lisp:53-56: (apply *solver* problem args)"""
    return SOLVER(problem, **kwargs)


# The generics (src/solver.lisp:59-80); the default methods are the ones on `tableau`.
@functools.singledispatch
def solution_problem(solution):
    raise TypeError(f"no solution-problem method for {type(solution).__name__}")


@functools.singledispatch
def solution_objective_value(solution):
    raise TypeError(f"no solution-objective-value method for {type(solution).__name__}")


@functools.singledispatch
def solution_variable(solution, variable):
    raise TypeError(f"no solution-variable method for {type(solution).__name__}")


@functools.singledispatch
def solution_reduced_cost(solution, variable):
    raise TypeError(f"no solution-reduced-cost method for {type(solution).__name__}")


solution_problem.register(simplex.Tableau, lambda s: s.problem)
solution_objective_value.register(simplex.Tableau, simplex.tableau_objective_value)
solution_variable.register(simplex.Tableau, simplex.tableau_variable)
solution_reduced_cost.register(simplex.Tableau, simplex.tableau_reduced_cost)


class SolutionVariables(dict):
    """What the body of with-solved-problem / with-solution-variables sees: the bound variables
    (item or attribute access) and the local `reduced-cost` macro (src/solver.lisp:101-103)."""

    def __init__(self, values, solution):
        super().__init__(values)
        self.solution = solution

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def reduced_cost(self, var):
        return solution_reduced_cost(self.solution, var)


def with_solution_variables(var_list, solution):
    """src/solver.lisp:96-115"""
    if isinstance(var_list, Problem):
        values = {var_list.objective_var: solution_objective_value(solution)}
        values.update({v: solution_variable(solution, v) for v in var_list.vars})
    else:
        values = {v: solution_variable(solution, v) for v in var_list}
    return SolutionVariables(values, solution)


def with_solved_problem(objective_func, *constraints, **kwargs):
    """src/solver.lisp:86-94: parse, solve through SOLVER, bind every variable."""
    problem = parse_linear_problem(objective_func, list(constraints))
    return with_solution_variables(problem, solve_problem(problem, **kwargs))


__all__ = ["SOLVER", "set_solver", "using_solver", "solve_problem", "solution_problem",
           "solution_objective_value", "solution_variable", "solution_reduced_cost",
           "with_solved_problem", "with_solution_variables", "SolutionVariables"]
