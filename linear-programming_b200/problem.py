"""The `problem` struct and the DSL parser (reference: src/problem.lisp).  Front end only: it
feeds build_tableau; nothing here is on the GPU path."""
import itertools
import warnings
from dataclasses import dataclass, field

from .conditions import ParsingError
from .expressions import (CONSTANT, parse_linear_expression, scale_linear_expression,
                          sum_linear_expressions)
from .sexp import as_form
from .utils import lb_max as _lb_max, ub_min as _ub_min, validate_bounds

_gensym = itertools.count()


class Uninterned(str):
    """An objective-variable name nobody wrote down: (gensym "Z") in parse-linear-problem
    (src/problem.lisp:167) or (make-symbol objective) in read-mps.  write-sexp leaves these out."""


@dataclass
class Problem:
    """src/problem.lisp:45-53.  constraints: [(op, [(var, coef)...], rhs)] with op in <=, >=, =;
    var_bounds: [(var, (lb, ub))] with None for infinity; objective_func: [(var, coef)]."""
    type: str = "max"
    vars: tuple = ()
    objective_var: str = "z"
    objective_func: list = field(default_factory=list)
    integer_vars: list = field(default_factory=list)
    var_bounds: list = field(default_factory=list)
    constraints: list = field(default_factory=list)



def _add_bound(table, var, new, implicit_lb=None):
    """src/problem.lisp:64-71"""
    if var in table:
        old = table[var]
        table[var] = (_lb_max(old[0], new[0]), _ub_min(old[1], new[1]))
    else:
        table[var] = (new[0] if new[0] is not None else implicit_lb, new[1])


def _is_number(x):
    from numbers import Number
    return isinstance(x, Number) and not isinstance(x, bool)


def parse_linear_constraints(exprs):
    """src/problem.lisp:73-156 -> (simple constraints, integer vars, bounds alist)."""
    bounds, equalities, integer = {}, [], []
    for expr in exprs:
        head = expr[0]
        if head in ("<=", "<"):
            if head == "<":
                warnings.warn("< constraints are deprecated in favor of <= ones due to misleading semantics.")
            equalities.append(("<=", [parse_linear_expression(e) for e in expr[1:]]))
        elif head in (">=", ">"):
            if head == ">":
                warnings.warn("> constraints are deprecated in favor of >= ones due to misleading semantics.")
            equalities.append(("<=", [parse_linear_expression(e) for e in expr[1:]][::-1]))
        elif head == "=":
            equalities.append(("=", [parse_linear_expression(e) for e in expr[1:]]))
        elif head == "integer":
            integer += [v for v in expr[1:] if v not in integer]
        elif head == "bounds":
            for entry in expr[1:]:
                if isinstance(entry[0], str):
                    if not (len(entry) <= 2 and (len(entry) < 2 or entry[1] is None or _is_number(entry[1]))):
                        raise ParsingError(f"Invalid bounds entry {entry!r}")
                    _add_bound(bounds, entry[0], (None, entry[1] if len(entry) > 1 else None))
                else:
                    ok = (_is_number(entry[0]) and len(entry) >= 2 and isinstance(entry[1], str)
                          and (len(entry) < 3 or entry[2] is None or _is_number(entry[2])))
                    if not ok:
                        raise ParsingError(f"Invalid bounds entry {entry!r}")
                    _add_bound(bounds, entry[1], (entry[0], entry[2] if len(entry) > 2 else None))
        elif head == "binary":
            integer += [v for v in expr[1:] if v not in integer]
            for var in expr[1:]:
                _add_bound(bounds, var, (0, 1))
        else:
            raise ParsingError(f"{expr!r} is not a valid constraint")

    simple = []
    for op, sides in equalities:
        for lhs, rhs in zip(sides, sides[1:]):
            lin = sum_linear_expressions(lhs, scale_linear_expression(rhs, -1))
            const = -lin.get(CONSTANT, 0)
            terms = [(v, c) for v, c in lin.items() if v != CONSTANT]
            if len(terms) == 1:                                    # single variable -> a bound
                var, coef = terms[0]
                c = const / coef if isinstance(const, float) or isinstance(coef, float) \
                    else _exact_div(const, coef)
                new = (c, c) if op == "=" else (c, None) if coef <= 0 else (None, c)
                _add_bound(bounds, var, new, 0)
            elif op == "=":
                simple.append(("=", terms, const))
            elif 0 <= const:
                simple.append(("<=", terms, const))
            else:
                simple.append((">=", [(v, -c) for v, c in terms], -const))
    for var, (lb, ub) in bounds.items():                           # validate-bounds, utils.lisp:69-76
        validate_bounds(lb, ub, var)
    return simple, integer, [(var, b) for var, b in bounds.items()]


def _exact_div(a, b):
    from fractions import Fraction
    q = Fraction(a) / Fraction(b)
    return int(q) if q.denominator == 1 else q


def parse_linear_problem(objective_exp, constraints):
    """src/problem.lisp:160-205.  Variable order = first appearance (objective, integer, bounds,
    constraints), a deterministic stand-in for the reference's hash-table order (:196-198)."""
    objective_exp = as_form(objective_exp)
    constraints = [as_form(c) for c in constraints]
    has_var = objective_exp[0] == "="
    objective = objective_exp[2] if has_var else objective_exp
    objective_var = objective_exp[1] if has_var else Uninterned(f"z{next(_gensym)}")
    if (not has_var and isinstance(objective[1], list) and objective[1] and objective[1][0] == "="):
        objective_var = objective[1][1]
        objective = [objective[0], objective[1][2]]
    if objective[0] not in ("min", "max"):
        raise ParsingError(f"{objective[0]} is neither min nor max in objective function {objective!r}")
    func = parse_linear_expression(objective[1])
    simple, integer, bounds = parse_linear_constraints(constraints)
    seen = {}
    for var in func:
        seen.setdefault(var, True)
    for var in integer:
        seen.setdefault(var, True)
    for var, _ in bounds:
        seen.setdefault(var, True)
    for _, terms, _ in simple:
        for var, _ in terms:
            seen.setdefault(var, True)
    return Problem(type=objective[0], vars=tuple(seen), objective_var=objective_var,
                   objective_func=list(func.items()), integer_vars=integer, var_bounds=bounds,
                   constraints=simple)


def make_linear_problem(objective, *constraints):
    """src/problem.lisp:208-210 (a macro there; here the forms are DSL text or nested lists)."""
    return parse_linear_problem(objective, list(constraints))
