"""Golden tableaus transcribed from the reference's own tests (exact rationals).

Source: /root/reference/t/simplex.lisp and README.md of
neil-lindquist/linear-programming @ 7fe5c78.  The reference is Common Lisp and cannot be
executed in this image, so these hand-transcribed known-answer vectors (each with its
file:line) are what pins the oracle and, through it, the CUDA path.  Column order is the
x, y, z order of the golden matrices themselves (the reference's `tableau-matrix-equal`,
t/simplex.lisp:15-31, permutes by variable name for the same reason).

Every entry: dict(name, source, is_max, initial tableau(s) + basis, expected tableau(s) +
basis, pivots).  Strings such as "7/2" are exact rationals (fractions.Fraction parses them).
"""
from fractions import Fraction as F


def M(rows):
    return [[F(x) for x in row] for row in rows]


# --- max x + 4y + 3z  s.t. 2x + y <= 8, y + z <= 7 -----------------------------------------
BASIC_INITIAL = dict(
    source="t/simplex.lisp:60-72",
    matrix=M([[2, 1, 0, 1, 0, 8], [0, 1, 1, 0, 1, 7], [-1, -4, -3, 0, 0, 0]]),
    basis=[3, 4], var_count=5, constraint_count=2, objective=0, is_max=True)

# one pivot on column x (0), row 0  -- t/simplex.lisp:135-159
SINGLE_PIVOT = dict(
    source="t/simplex.lisp:135-159",
    initial=BASIC_INITIAL, col=0, row=0,
    matrix=M([[1, "1/2", 0, "1/2", 0, 4], [0, 1, 1, 0, 1, 7], [0, "-7/2", -3, "1/2", 0, 4]]),
    basis=[0, 4], objective=4)

# full solve -- t/simplex.lisp:170-194, README.md:58-62 (obj 57/2, x=1/2, y=7, z=0)
BASIC_SOLVED = dict(
    source="t/simplex.lisp:170-194",
    initial=BASIC_INITIAL,
    matrix=M([[1, 0, "-1/2", "1/2", "-1/2", "1/2"], [0, 1, 1, 0, 1, 7],
              [0, 0, "1/2", "1/2", "7/2", "57/2"]]),
    basis=[0, 1], objective=F(57, 2), pivots=2,
    trace=[(1, 1), (0, 0)],  # Dantzig picks y (-4) first, then x
    primal=dict(x=F(1, 2), y=7, z=0), reduced_cost=dict(x=0, y=0, z=F(1, 2)))

# --- + (= (+ (* 2 x) y z) 8): two-phase with an equality row -- t/simplex.lisp:74-103 ------
EQ_BUILD = dict(
    source="t/simplex.lisp:74-103",
    art_matrix=M([[2, 1, 0, 1, 0, 0, 8], [0, 1, 1, 0, 1, 0, 7], [2, 1, 1, 0, 0, 1, 8],
                  [2, 1, 1, 0, 0, 0, 8]]),
    art_basis=[3, 4, 5], art_var_count=6, art_objective=8,
    main_matrix=M([[2, 1, 0, 1, 0, 8], [0, 1, 1, 0, 1, 7], [2, 1, 1, 0, 0, 8],
                   [-1, -4, -3, 0, 0, 0]]),
    main_basis=[3, 4, 6], main_var_count=5, constraint_count=3, is_max=True)

EQ_SOLVED = dict(  # t/simplex.lisp:196-237 ; phase 1 has a ratio tie rows 0/2 -> row 0 wins
    source="t/simplex.lisp:196-237",
    initial=EQ_BUILD,
    art_matrix=M([[1, "1/2", 0, "1/2", 0, 0, 4], [0, 1, 0, 1, 1, -1, 7], [0, 0, 1, -1, 0, 1, 0],
                  [0, 0, 0, 0, 0, -1, 0]]),
    art_basis=[0, 4, 2], art_objective=0,
    main_matrix=M([[1, 0, 0, 0, "-1/2", "1/2"], [0, 1, 0, 1, 1, 7], [0, 0, 1, -1, 0, 0],
                   [0, 0, 0, 1, "7/2", "57/2"]]),
    main_basis=[0, 1, 2], objective=F(57, 2), pivots=(2, 0, 1))

# --- + (>= (+ x z) 1): two-phase with a surplus column -- t/simplex.lisp:105-133 -----------
GEQ_BUILD = dict(
    source="t/simplex.lisp:105-133",
    art_matrix=M([[2, 1, 0, 1, 0, 0, 0, 8], [0, 1, 1, 0, 1, 0, 0, 7], [1, 0, 1, 0, 0, -1, 1, 1],
                  [1, 0, 1, 0, 0, -1, 0, 1]]),
    art_basis=[3, 4, 6], art_var_count=7, art_objective=1,
    main_matrix=M([[2, 1, 0, 1, 0, 0, 8], [0, 1, 1, 0, 1, 0, 7], [1, 0, 1, 0, 0, -1, 1],
                   [-1, -4, -3, 0, 0, 0, 0]]),
    main_basis=[3, 4, 7], main_var_count=6, constraint_count=3, is_max=True)

GEQ_SOLVED = dict(  # t/simplex.lisp:239-275, first accepted alternative (var order x, y, z)
    source="t/simplex.lisp:239-275",
    initial=GEQ_BUILD,
    art_matrix=M([[0, 1, -2, 1, 0, 2, -2, 6], [0, 1, 1, 0, 1, 0, 0, 7], [1, 0, 1, 0, 0, -1, 1, 1],
                  [0, 0, 0, 0, 0, 0, -1, 0]]),
    art_basis=[3, 4, 0], art_objective=0,
    main_matrix=M([[0, 1, 0, "1/3", "2/3", "2/3", "20/3"], [0, 0, 1, "-1/3", "1/3", "-2/3", "1/3"],
                   [1, 0, 0, "1/3", "-1/3", "-1/3", "2/3"],
                   [0, 0, 0, "2/3", "10/3", "1/3", "85/3"]]),
    main_basis=[1, 2, 0], objective=F(85, 3), pivots=(1, 0, 2))

# --- infeasible / unbounded -- t/simplex.lisp:277-289 ----------------------------------------
# max x+4y+3z s.t. 2x+y<=4, y+z<=2, x+z>=5  -> infeasible-problem-error
INFEASIBLE = dict(
    source="t/simplex.lisp:278-283",
    art_matrix=M([[2, 1, 0, 1, 0, 0, 0, 4], [0, 1, 1, 0, 1, 0, 0, 2], [1, 0, 1, 0, 0, -1, 1, 5],
                  [1, 0, 1, 0, 0, -1, 0, 5]]),
    art_basis=[3, 4, 6],
    main_matrix=M([[2, 1, 0, 1, 0, 0, 4], [0, 1, 1, 0, 1, 0, 2], [1, 0, 1, 0, 0, -1, 5],
                   [-1, -4, -3, 0, 0, 0, 0]]),
    main_basis=[3, 4, 7], is_max=True)
# max x+4y+3z s.t. 2x+y<=4, y-z<=4 -> unbounded-problem-error
UNBOUNDED = dict(
    source="t/simplex.lisp:285-289",
    matrix=M([[2, 1, 0, 1, 0, 4], [0, 1, -1, 0, 1, 4], [-1, -4, -3, 0, 0, 0]]),
    basis=[3, 4], is_max=True)

# --- assembly LP, t/integration.lisp:32-58 (columns: widgets d1 d2 d3 | 4 slacks | rhs) ----
ASSEMBLY = dict(
    source="t/integration.lisp:32-58",
    matrix=M([[4, -7, -6, -8, 1, 0, 0, 0, 0], [3, -5, -9, -4, 0, 1, 0, 0, 0],
              [0, 8, 5, 3, 0, 0, 1, 0, 100], [0, 6, 9, 8, 0, 0, 0, 1, 200],
              [-3, 0, 0, 0, 0, 0, 0, 0, 0]]),
    basis=[4, 5, 6, 7], is_max=True, pivots=4,
    objective=F(160200, 1177),
    primal=dict(widgets=F(53400, 1177), d1=F(2800, 1177), d2=F(8200, 1177), d3=F(18100, 1177)),
    bounds=dict(revenue=(136.08, 136.11), widgets=(45.36, 45.37), d1=(2.37, 2.38),
                d2=(6.96, 6.97), d3=(15.37, 15.38)))

# --- Beale's cycling example (not from the reference; textbook, Chvatal p.31 variant) ------
# max 3/4 x1 - 150 x2 + 1/50 x3 - 6 x4  (written as integers x4 to stay exact in fp64)
# Dantzig + lowest-index ties cycles with period 6; Bland's rule terminates (obj 1/20).
BEALE = dict(
    source="textbook (Beale 1955); build extension a5, not in the reference",
    matrix=M([["1/4", -60, "-1/25", 9, 1, 0, 0, 0],
              ["1/2", -90, "-1/50", 3, 0, 1, 0, 0],
              [0, 0, 1, 0, 0, 0, 1, 1],
              ["-3/4", 150, "-1/50", 6, 0, 0, 0, 0]]),
    basis=[4, 5, 6], is_max=True, objective=F(1, 20))
