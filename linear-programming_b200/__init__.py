"""linear_programming_b200 -- B200-native dense simplex backend for the `*solver*` hook of
neil-lindquist/linear-programming (reference @ 7fe5c78).

csrc/         CUDA kernels + the C ABI (include/b200lp.h) -> libb200lp.so
_ffi.py       ctypes binding of that ABI (what the Lisp CFFI shim binds too)
lisp/         the CFFI shim installing `b200-solver` into `*solver*` (for Lisp hosts)
solver.py     the hook itself as the reference exposes it (src/solver.lisp): SOLVER, solve_problem,
              solution_* generics, with_solved_problem -- default backend = simplex.b200_solver
simplex.py    tableau struct, build_tableau, accessors, branch and bound (src/simplex.lisp minus
              the hot path, which only exists on the GPU)
problem.py / expressions.py / sexp.py / conditions.py   the DSL front end feeding build_tableau
external_formats.py   read_sexp / write_sexp / read_mps / write_standard_format
utils.py      bound helpers and the fp=/fp</fp> tolerance predicates the device thresholds restate
sharded.py    one-process-per-GPU row-block sharding (host plumbing)
synthetic.py  BASELINE.json's synthetic dense LPs
"""
from . import _ffi  # noqa: F401
from ._ffi import (B200DeviceError, B200LibraryError, DeviceTableau, make_opts)  # noqa: F401
from .conditions import (InfeasibleProblemError, ParsingError, SolverError,  # noqa: F401
                         UnboundedProblemError)
from .external_formats import read_mps, read_sexp, write_sexp, write_standard_format  # noqa: F401
from .problem import Problem, make_linear_problem, parse_linear_problem  # noqa: F401
from .simplex import Tableau, b200_solver, build_tableau, solve_tableau  # noqa: F401
from .solver import (solution_objective_value, solution_problem, solution_reduced_cost,  # noqa: F401
                     solution_variable, solve_problem, using_solver, with_solution_variables,
                     with_solved_problem)
