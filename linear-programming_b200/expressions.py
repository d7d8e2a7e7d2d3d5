"""Linear-expression algebra of the DSL front end (reference: src/expressions.lisp).  An
expression is an insertion-ordered dict {variable: coefficient}; constants live under CONSTANT."""
from fractions import Fraction
from numbers import Number

from .conditions import NonlinearError, ParsingError

CONSTANT = "+constant+"


def _is_constant(expr):
    return len(expr) == 1 and CONSTANT in expr                       # linear-constant-p :21-24


def sum_linear_expressions(*exprs):
    """src/expressions.lisp:27-33"""
    total = dict(exprs[0])
    for e in exprs[1:]:
        for var, coef in e.items():
            total[var] = total.get(var, 0) + coef
    return total


def scale_linear_expression(expr, scalar):
    """src/expressions.lisp:37-40"""
    return {var: scalar * coef for var, coef in expr.items()}


def _recip(x):
    return Fraction(1) / x if isinstance(x, (int, Fraction)) else 1.0 / x


def parse_linear_expression(expr):
    """src/expressions.lisp:43-108: symbol | number | (+ - * / ...) | (:alist ...) | (:plist ...)."""
    if isinstance(expr, str):
        return {expr: 1}
    if isinstance(expr, Number) and not isinstance(expr, bool):
        return {CONSTANT: expr}
    if isinstance(expr, (list, tuple)) and expr:
        head, args = expr[0], list(expr[1:])
        if head == ":alist":
            return {pair[0]: pair[-1] for pair in args}
        if head == ":plist":
            return {args[i]: args[i + 1] for i in range(0, len(args), 2)}
        if head == "+":
            return sum_linear_expressions(*[parse_linear_expression(a) for a in args])
        if head == "*":
            variable, constant = None, 1
            for fact in (parse_linear_expression(a) for a in args):
                if _is_constant(fact):
                    constant = constant * fact[CONSTANT]
                elif variable is not None:
                    raise NonlinearError(expr)
                else:
                    variable = fact
            return scale_linear_expression(variable, constant) if variable is not None \
                else {CONSTANT: constant}
        if head == "-":
            first = parse_linear_expression(args[0])
            if len(args) == 1:
                return scale_linear_expression(first, -1)
            rest = parse_linear_expression(["+"] + args[1:])
            return sum_linear_expressions(first, scale_linear_expression(rest, -1))
        if head == "/":
            first = parse_linear_expression(args[0])
            if len(args) == 1:
                if not _is_constant(first):
                    raise NonlinearError(expr)
                return {CONSTANT: _recip(first[CONSTANT])}
            divisors = [parse_linear_expression(a) for a in args[1:]]
            if not all(_is_constant(d) for d in divisors):
                raise NonlinearError(expr)
            prod = 1
            for d in divisors:
                prod = prod * d[CONSTANT]
            return scale_linear_expression(first, _recip(prod))
        raise NonlinearError(expr)
    raise ParsingError(f"{expr!r} is not a symbol, number, or an expression")


def format_linear_expression(expr):
    """src/expressions.lisp:111-118"""
    return ["+"] + [coef if var == CONSTANT else ["*", coef, var] for var, coef in expr.items()]
