"""Reading and writing problems in external formats (reference: src/external-formats.lisp) --
the loader that lets benchmarks and users feed standard LP files to the B200 backend
(SURVEY.md 8 f4).  Host-side front end only: O(file size), nothing here touches the GPU.

  read_sexp / write_sexp        src/external-formats.lisp:44-76
  read_mps                      :78-348   fixed-width MPS: ROWS/COLUMNS/RHS/RANGES/BOUNDS plus the
                                          OBJSENSE / OBJNAME extensions and BV / LI / UI bounds
  write_standard_format         :350-405

Names: this package's DSL reader folds symbols to lower case (the stand-in for the Lisp reader's
up-casing), so read_mps's default `read_case` is "downcase"; "upcase", "preserve" and "invert"
behave as in the reference.  Deviations from the reference, all in rarely used corners where the
reference code cannot work as written, are marked DEVIATION below.
"""
import io
from fractions import Fraction

from .conditions import InvalidBoundsError, ParsingError
from .expressions import format_linear_expression
from .problem import Problem, Uninterned, parse_linear_problem
from . import sexp


# ------------------------------------------------------------------------------------------ sexp
def _read_one_form_text(stream):
    """Consume exactly one parenthesised form from a text stream (the rest stays unread,
    t/external-formats.lisp:104-112)."""
    out, depth, started, in_str = [], 0, False, False
    while True:
        ch = stream.read(1)
        if ch == "":
            break
        out.append(ch)
        if in_str:
            if ch == "\\":
                out.append(stream.read(1))
            elif ch == '"':
                in_str = False
            continue
        if ch == '"':
            in_str = True
        elif ch == ";":
            while ch not in ("", "\n"):
                ch = stream.read(1)
            out[-1] = "\n"
        elif ch == "(":
            depth += 1
            started = True
        elif ch == ")":
            depth -= 1
            if started and depth == 0:
                break
    return "".join(out)


def _eval_read_time(form):
    """#.(...) support for trusted input: arithmetic on numbers only."""
    if not isinstance(form, list):
        return form
    op, args = form[0], [_eval_read_time(a) for a in form[1:]]
    if op == "+":
        return sum(args)
    if op == "*":
        r = 1
        for a in args:
            r = r * a
        return r
    if op == "-":
        return -args[0] if len(args) == 1 else args[0] - sum(args[1:])
    if op == "/":
        r = Fraction(args[0]) if isinstance(args[0], int) else args[0]
        for a in args[1:]:
            r = r / a
        return int(r) if isinstance(r, Fraction) and r.denominator == 1 else r
    raise ParsingError(f"cannot evaluate #.{form!r} at read time")


def _resolve_read_eval(form, allow):
    if isinstance(form, list):
        out, i = [], 0
        while i < len(form):
            if form[i] == "#." and i + 1 < len(form):
                if not allow:
                    raise ParsingError("#. read-time evaluation is disabled (allow_read_eval=False)")
                out.append(_eval_read_time(form[i + 1]))
                i += 2
            else:
                out.append(_resolve_read_eval(form[i], allow))
                i += 1
        return out
    return form


def _strip_package(form):
    """`linear-programming/problem:bounds` -> `bounds` (t/external-formats.lisp:46)."""
    if isinstance(form, str) and ":" in form and not form.startswith(":"):
        return form.rsplit(":", 1)[1]
    if isinstance(form, list):
        return [_strip_package(f) for f in form]
    return form


def read_sexp(stream, allow_read_eval=False, package=None):
    """src/external-formats.lisp:44-54: one sexp, first element the objective, the rest the
    constraints.  `stream` is a text stream or a string."""
    if isinstance(stream, str):
        stream = io.StringIO(stream)
    text = _read_one_form_text(stream).replace("#.", " #. ")
    form = _strip_package(_resolve_read_eval(sexp.read(text), allow_read_eval))
    return parse_linear_problem(form[0], form[1:])


def _fmt_number(x):
    if x is None:
        return "nil"
    if isinstance(x, Fraction):
        return f"{x.numerator}/{x.denominator}" if x.denominator != 1 else str(x.numerator)
    if isinstance(x, float):
        return repr(x).replace("e", "d") if "e" in repr(x) else repr(x) + "d0"
    return str(x)


def _fmt_form(form):
    if isinstance(form, list):
        return "(" + " ".join(_fmt_form(f) for f in form) + ")"
    if isinstance(form, str):
        return form
    return _fmt_number(form)


def write_sexp(stream, problem, package=None):
    """src/external-formats.lisp:56-76.  DEVIATION: bounds are written in the DSL's own
    `(lb var ub)` / `(var ub)` grammar (docs/linear-problem-syntax.md) so that every problem
    round-trips; the reference prints its internal dotted pairs, which read-sexp only accepts for
    free variables."""
    objective = [problem.type, format_linear_expression(dict(problem.objective_func))]
    if not isinstance(problem.objective_var, Uninterned):
        objective = ["=", problem.objective_var, objective]       # user-named objective
    forms = [objective]
    if problem.integer_vars:
        forms.append(["integer"] + list(problem.integer_vars))
    if problem.var_bounds:
        entries = []
        for var, (lb, ub) in problem.var_bounds:
            entries.append([var] + ([ub] if ub is not None else []) if lb is None else [lb, var, ub])
        forms.append(["bounds"] + entries)
    for op, terms, rhs in problem.constraints:
        forms.append([op, format_linear_expression(dict(terms)), rhs])
    stream.write(_fmt_form(forms) + "\n")


# ------------------------------------------------------------------------------------------- MPS
_FIELD_START = (0, 1, 4, 14, 24, 39, 49)      # src/external-formats.lisp:101-102 (field 0 = whole line)
_FIELD_END = (61, 3, 12, 22, 36, 47, 61)


def _parse_number(raw, number_type):
    """:129-165.  Decimal text -> exact rational (default) or float.  DEVIATION: exponents are
    parsed (the reference's exponent branch overwrites the mantissa and cannot work)."""
    raw = raw.strip()
    if number_type in ("rational", Fraction):
        text = raw.lower().replace("d", "e")
        value = Fraction(text)
        return int(value) if value.denominator == 1 else value
    return float(raw.lower().replace("d", "e"))


def read_mps(stream, problem_type=None, package=None, read_case="downcase", trim_names_p=True,
             number_type="rational", rhs_id=None):
    """src/external-formats.lisp:78-348.  A line starting with ENDATA ends the problem, so MPS
    data can be embedded in a longer stream."""
    if isinstance(stream, str):
        stream = io.StringIO(stream)

    def casefold(raw):
        if read_case == "upcase":
            return raw.upper()
        if read_case == "downcase":
            return raw.lower()
        if read_case == "preserve":
            return raw
        if read_case == "invert":
            if raw and all(c.isupper() for c in raw):
                return raw.lower()
            if raw and all(c.islower() for c in raw):
                return raw.upper()
            return raw
        raise ValueError(f"unknown read_case {read_case!r}")

    def field(n, line, kind="raw"):
        raw = line[min(len(line), _FIELD_START[n]):min(len(line), _FIELD_END[n])]
        if kind == "raw":
            return raw
        if kind == "number":
            return _parse_number(raw, number_type)
        return casefold(raw.strip(" ") if trim_names_p else raw)        # symbol / name-string

    rows = {}            # row name -> [type, rhs, range, terms(list, newest first)]
    var_info = {}        # var -> [lb, ub, integer?]
    objective = None
    header = None
    for raw_line in stream:
        line = raw_line.rstrip("\n").rstrip(" \r")
        if not line:
            continue
        try:
            header, objective, problem_type, rhs_id = _mps_record(
                line, header, objective, problem_type, rhs_id, rows, var_info, field)
        except ParsingError:
            raise
        except KeyError as exc:                           # a record names a row ROWS never declared
            raise ParsingError(f"unknown row {exc.args[0]!r} in MPS record {line!r}") from exc
        except (ValueError, ZeroDivisionError) as exc:    # an empty or malformed numeric field
            raise ParsingError(f"bad number in MPS record {line!r}: {exc}") from exc
        if header == "endata":
            break
    return _mps_finish(rows, var_info, objective, problem_type, number_type)


def _mps_record(line, header, objective, problem_type, rhs_id, rows, var_info, field):
    """One line of the file; returns the updated (header, objective, problem_type, rhs_id)."""
    if True:
        if line[0] != " ":
            card = line[:15].lower()
            if card[0] == "*":
                return header, objective, problem_type, rhs_id
            if card == "endata":
                return "endata", objective, problem_type, rhs_id
            return card.split()[0], objective, problem_type, rhs_id   # NAME carries no body records
        if header == "rows":
            kind = {"n": "objective", "g": ">=", "l": "<=", "e": "="}.get(field(1, line)[:1].lower())
            name = field(2, line, "name")
            if kind == "objective" and objective is None:
                objective = name                          # the first N row is the objective
            rows[name] = [kind, 0, None, []]
        elif header == "columns":
            var = field(2, line, "name")
            var_info.setdefault(var, [0, None, False])
            rows[field(3, line, "name")][3].insert(0, (var, field(4, line, "number")))
            if len(field(5, line)) != 0:
                rows[field(5, line, "name")][3].insert(0, (var, field(6, line, "number")))
        elif header == "rhs":
            # (string= rhs-id current-rhs-id), :215-218: the caller's :rhs-id is compared as given
            # (no case folding) with the set name as the file's read-case leaves it.  An RHS entry
            # on the OBJECTIVE row (by MPS convention minus the objective's constant) is stored and
            # then ignored, as in the reference (:282-285 skips the objective row).
            current = field(2, line, "name")
            if rhs_id is None:
                rhs_id = current
            if rhs_id == current:
                rows[field(3, line, "name")][1] = field(4, line, "number")
                if len(field(5, line)) != 0:
                    rows[field(5, line, "name")][1] = field(6, line, "number")
        elif header == "ranges":
            # DEVIATION: rows are looked up by name (the reference interns the name as a symbol
            # and then misses its own string-keyed table, :229-237)
            rows[field(3, line, "name")][2] = field(4, line, "number")
            if len(field(5, line)) != 0:
                rows[field(5, line, "name")][2] = field(6, line, "number")
        elif header == "bounds":
            var = field(3, line, "name")
            lb, ub, intp = var_info.get(var, [0, None, False])
            kind = field(1, line, "name").upper()
            if kind == "LO":
                lb = field(4, line, "number")
            elif kind == "UP":
                ub = field(4, line, "number")
            elif kind == "FX":
                lb = ub = field(4, line, "number")
            elif kind == "FR":
                lb, ub = None, None
            elif kind == "MI":
                lb = None
            elif kind == "PL":
                ub = None
            elif kind == "BV":
                lb, ub, intp = 0, 1, True
            elif kind == "LI":
                lb, intp = field(4, line, "number"), True
            elif kind == "UI":
                ub, intp = field(4, line, "number"), True
            else:
                raise ParsingError(f"{kind!r} is not a know bound type")
            var_info[var] = [lb, ub, intp]
        elif header == "objsense":
            header = None                                 # only one record for this header
            kind = field(0, line, "name").lower()
            if kind in ("max", "maximizing"):
                problem_type = "max"
            elif kind in ("min", "minimizing"):
                problem_type = "min"
            else:
                raise ParsingError(f"{kind!r} is not a know problem type")
        elif header == "objname":
            header = None
            objective = field(0, line, "name")
        else:
            raise ParsingError(f"Unknown header-card {header}")
        return header, objective, problem_type, rhs_id


def _mps_finish(rows, var_info, objective, problem_type, number_type):
    """src/external-formats.lisp:282-348: rows -> constraints, single-variable rows -> bounds."""
    if problem_type not in ("max", "min"):
        raise ParsingError("No valid problem type was specified")

    constraints = []
    for name, (op, rhs, rng, terms) in rows.items():      # :282-300
        if op == "objective":
            continue
        constraints.insert(0, [op, list(terms), rhs])
        if rng is not None:
            if op == "<=":
                constraints.insert(0, [">=", list(terms), rhs - abs(rng)])
            elif op == ">=":
                constraints.insert(0, ["<=", list(terms), rhs + abs(rng)])
            elif rng > 0:
                constraints.insert(0, ["<=", list(terms), rhs + rng])
            elif rng < 0:
                constraints.insert(0, [">=", list(terms), rhs + rng])
    kept = []
    for op, terms, rhs in constraints:                    # :301-324
        if len(terms) == 1:
            # DEVIATION: a single-variable row tightens that variable's bound (<= the upper, >=
            # the lower bound, respecting the coefficient's sign); the reference writes the
            # wrong slots of its (lb ub integer-p) record here
            var, coef = terms[0]
            if coef == 0:
                raise ParsingError(f"row with the single variable {var!r} has a zero coefficient")
            bound = Fraction(rhs) / Fraction(coef) if number_type in ("rational", Fraction) else rhs / coef
            if isinstance(bound, Fraction) and bound.denominator == 1:
                bound = int(bound)
            info = var_info[var]
            upper = (op == "<=") == (coef > 0)
            if op == "=" or upper:
                info[1] = bound if info[1] is None else min(info[1], bound)
            if op == "=" or not upper:
                info[0] = bound if info[0] is None else max(info[0], bound)
        elif rhs < 0:
            flipped = {"<=": ">=", ">=": "<=", "=": "="}[op]
            kept.append((flipped, [(v, -c) for v, c in terms], -rhs))
        else:
            kept.append((op, terms, rhs))
    int_vars, bounds = [], []
    for var, (lb, ub, intp) in var_info.items():          # :325-333
        if intp:
            int_vars.insert(0, var)
        if lb != 0 or lb is None or ub is not None:
            if lb is not None and ub is not None and ub < lb:
                raise InvalidBoundsError(var, ub, lb)
            bounds.insert(0, (var, (lb, ub)))
    return Problem(type=problem_type, vars=tuple(var_info), objective_var=Uninterned(objective),
                   objective_func=list(rows[objective][3]) if objective in rows else [],
                   integer_vars=int_vars, var_bounds=bounds, constraints=kept)


# ------------------------------------------------------------------------- human readable writer
def _print_linear_expression(expression):
    """src/external-formats.lisp:350-358 (kept as written there, sign convention included)."""
    out = []
    for k, (var, coef) in enumerate(expression):
        if k == 0:
            if 0 < coef:
                out.append("-")
        else:
            out.append(" - " if 0 < coef else " + ")
        if coef not in (1, -1):
            out.append(f"{_fmt_plain(abs(coef))}*")
        out.append(str(var))
    return "".join(out)


def _fmt_plain(x):
    if isinstance(x, Fraction):
        return f"{x.numerator}/{x.denominator}" if x.denominator != 1 else str(x.numerator)
    return str(x)


def write_standard_format(stream, problem, unicodep=True, aesthetic_variable_names_p=True):
    """src/external-formats.lisp:360-405"""
    le, ge = ("≤", "≥") if unicodep else ("<", ">")
    pad = " " * 12
    stream.write(f"{'Maximize' if problem.type == 'max' else 'Minimize'} {problem.objective_var} = ")
    stream.write(_print_linear_expression(problem.objective_func))
    stream.write("\nSubject to:")
    for k, (op, terms, rhs) in enumerate(problem.constraints):
        stream.write((" " if k == 0 else pad) + _print_linear_expression(terms))
        stream.write(f" {dict([('<=', le), ('>=', ge), ('=', '=')])[op]} {_fmt_plain(rhs)}\n")
    bounds = dict(problem.var_bounds)
    non_negative = []
    for var in problem.vars:
        lb, ub = bounds.get(var, (0, None))
        if lb is not None:
            if lb == 0:
                non_negative.insert(0, var)
            else:
                stream.write(f"{pad}{var} {ge} {_fmt_plain(lb)}\n")
        if ub is not None:
            stream.write(f"{pad}{var} {le} {_fmt_plain(ub)}\n")
    if non_negative:
        stream.write(f"{pad}{', '.join(map(str, non_negative))} {ge} 0\n")
    if problem.integer_vars:
        stream.write(f"{pad}{', '.join(map(str, problem.integer_vars))} integer\n")


__all__ = ["read_sexp", "write_sexp", "read_mps", "write_standard_format"]
