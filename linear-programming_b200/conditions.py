"""Error hierarchy of the reference (src/conditions.lisp:15-77) and the mapping from the C ABI's
status codes onto it (include/b200lp.h)."""


class ParsingError(Exception):
    """parsing-error, src/conditions.lisp:15-20."""

    def __init__(self, description=""):
        self.description = description
        super().__init__(description)


class NonlinearError(ParsingError):
    """nonlinear-error, :22-30."""

    def __init__(self, expression):
        self.expression = expression
        super().__init__(f"{expression!r} is not a linear expression")


class InvalidBoundsError(ParsingError):
    """invalid-bounds-error, :32-41."""

    def __init__(self, var, ub, lb):
        self.var, self.ub, self.lb = var, ub, lb
        super().__init__(f"The bounds for variable {var} are invalid. Upper bound={ub}, Lower bound={lb}")


class SolverError(Exception):
    """solver-error, :43-45."""


class UnboundedProblemError(SolverError):
    """unbounded-problem-error, :47-53."""

    def __init__(self):
        super().__init__("Problem is unbounded")


class InfeasibleProblemError(SolverError):
    """infeasible-problem-error, :55-60."""

    def __init__(self, msg="Problem has no feasible region"):
        super().__init__(msg)


class InfeasibleIntegerConstraintsError(InfeasibleProblemError):
    """infeasible-integer-constraints-error, :62-67."""

    def __init__(self):
        super().__init__("Integer constrains could not be satisfied")


class UnsupportedConstraintError(SolverError):
    """unsupported-constraint-error, :69-77."""

    def __init__(self, constraint, solver_name):
        self.constraint, self.solver_name = constraint, solver_name
        super().__init__(f"{constraint!r} cannot be handled by the {solver_name} solver")


def raise_for_status(status):
    """C status -> condition (include/b200lp.h): 1 unbounded, 2 infeasible, others solver-error."""
    if status == 0:
        return
    if status == 1:
        raise UnboundedProblemError()
    if status == 2:
        raise InfeasibleProblemError()
    if status == 3:
        raise SolverError("iteration limit reached")
    if status == 4:
        raise SolverError("Artificial variable still in basis and cannot be replaced")
    if status == 5:
        raise SolverError("Artificial variable still non-zero")
    raise SolverError(f"b200lp status {status}")
