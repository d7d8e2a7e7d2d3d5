"""A minimal s-expression reader so problems can be written exactly as in the reference's DSL
(docs/linear-problem-syntax.md) and its tests.  Symbols become lower-case `str`, integers `int`,
ratios `fractions.Fraction`, `nil` None.  Like the Common Lisp reader with the default
*read-default-float-format*, a plain decimal such as 0.6861807 is a SINGLE-float: it is rounded
to float32 and then widened exactly, which is the value the reference's solver sees
(t/integration.lisp:74-80); a `d` exponent marker reads a double."""
import re
from fractions import Fraction

import numpy as np

_TOKEN = re.compile(r"""\s*(;[^\n]*|[()']|"(?:[^"\\]|\\.)*"|[^\s()';]+)""")
_INT = re.compile(r"^[+-]?\d+\.?$")
_RATIO = re.compile(r"^[+-]?\d+/\d+$")
_FLOAT = re.compile(r"^[+-]?(\d+\.\d*|\.\d+|\d+)([esfdlESFDL][+-]?\d+)?$")


def atom(tok):
    if _INT.match(tok):
        return int(tok.rstrip("."))
    if _RATIO.match(tok):
        return Fraction(tok)
    m = _FLOAT.match(tok)
    if m and any(ch in tok for ch in ".eEsSfFdDlL"):
        exp = m.group(2)
        marker = exp[0].lower() if exp else "e"
        text = tok if not exp else tok[:m.start(2)] + "e" + exp[1:]
        if marker in "dl":
            return float(text)
        return float(np.float32(text))
    low = tok.lower()
    return None if low == "nil" else low


def read_all(text):
    tokens = [t for t in _TOKEN.findall(text) if not t.startswith(";")]
    pos = 0

    def parse():
        nonlocal pos
        tok = tokens[pos]
        pos += 1
        if tok == "'":
            return parse()
        if tok == "(":
            out = []
            while tokens[pos] != ")":
                out.append(parse())
            pos += 1
            return out
        if tok == ")":
            raise SyntaxError("unbalanced )")
        return atom(tok)

    forms = []
    while pos < len(tokens):
        forms.append(parse())
    return forms


def read(text):
    forms = read_all(text)
    if len(forms) != 1:
        raise SyntaxError(f"expected one form, got {len(forms)}")
    return forms[0]


def as_form(x):
    """Accept DSL text or an already nested list/tuple structure."""
    if isinstance(x, str) and (x.lstrip().startswith("(") or x.lstrip().startswith("'")):
        return read(x)
    if isinstance(x, tuple):
        return [as_form(y) for y in x]
    if isinstance(x, list):
        return [as_form(y) for y in x]
    return x
