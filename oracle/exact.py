"""Exact-rational restatement of the reference simplex hot path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE -- only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module.

The reference's tableau is a ``(simple-array real 2)`` (src/simplex.lisp:53): with
integer/ratio inputs the whole solve runs in exact rational arithmetic, which is why
its tests compare with ``=`` against 57/2, 85/3 ... (t/simplex.lisp:190, 267).  This
module restates the same rules over ``fractions.Fraction`` so that those golden
tableaus can be reproduced bit-for-bit; it is what pins the fp64 C oracle
(oracle/simplex_oracle.c).  Pure-Python loops: small cases only.

Every function cites the reference lines it follows.  A tableau here is a list of
row lists (R x C), last column = RHS, last row = objective (src/simplex.lisp:74-78).
"""
from fractions import Fraction

OPTIMAL, UNBOUNDED, INFEASIBLE, ITERATION_LIMIT, ARTIFICIAL_STUCK = 0, 1, 2, 3, 4


def to_fraction_matrix(rows):
    return [[Fraction(x) for x in row] for row in rows]


def find_entering_column(tab, is_max, rule=0):
    """src/simplex.lisp:362-379 -- rationals compare exactly (src/utils.lisp:104)."""
    obj = tab[-1]
    nv = len(obj) - 1
    if nv <= 0:
        return None
    if rule == 1:  # Bland (build extension)
        for i in range(nv):
            if (obj[i] < 0) if is_max else (obj[i] > 0):
                return i
        return None
    best = 0
    for i in range(1, nv):
        if (obj[i] < obj[best]) if is_max else (obj[i] > obj[best]):
            best = i
    if (obj[best] < 0) if is_max else (obj[best] > 0):
        return best
    return None


def find_pivoting_row(tab, basis, j, rule=0):
    """src/simplex.lisp:382-389 -- first argmin of rhs/a over rows with a > 0."""
    m = len(tab) - 1
    best_row, best_q = None, None
    for i in range(m):
        a = tab[i][j]
        if a > 0:
            q = tab[i][-1] / a
            if (best_row is None or q < best_q or
                    (rule == 1 and q == best_q and basis[i] < basis[best_row])):
                best_row, best_q = i, q
    return best_row


def n_pivot_row(tab, basis, j, p):
    """src/simplex.lisp:337-359."""
    s = tab[p][j]
    tab[p] = [x / s for x in tab[p]]
    prow = tab[p]
    for r in range(len(tab)):
        if r == p:
            continue
        scale = tab[r][j]
        tab[r] = [x - scale * y for x, y in zip(tab[r], prow)]
    basis[p] = j


def n_solve_tableau(tab, basis, is_max, rule=0, max_iters=0, trace=None):
    """Single-tableau branch, src/simplex.lisp:453-461. Returns (status, pivots)."""
    it = 0
    while True:
        j = find_entering_column(tab, is_max, rule)
        if j is None:
            return OPTIMAL, it
        if max_iters and it >= max_iters:
            return ITERATION_LIMIT, it
        p = find_pivoting_row(tab, basis, j, rule)
        if p is None:
            return UNBOUNDED, it
        if trace is not None:
            trace.append((j, p))
        n_pivot_row(tab, basis, j, p)
        it += 1


def n_solve_two_phase(art, art_basis, main, main_basis, is_max, rule=0):
    """List branch, src/simplex.lisp:402-452. Returns (status, (it1, fix, it2))."""
    m = len(main) - 1
    num_vars = len(main[0]) - 1
    num_art_vars = len(art[0]) - 1
    st, it1 = n_solve_tableau(art, art_basis, False, rule)
    if st != OPTIMAL:
        return st, (it1, 0, 0)
    if art[m][num_art_vars] != 0:                                  # :405-407
        return INFEASIBLE, (it1, 0, 0)
    fix = 0
    for i in range(m):                                             # :419-434
        if art_basis[i] >= num_vars:
            if art[i][num_art_vars] != 0:
                return ARTIFICIAL_STUCK, (it1, fix, 0)
            new_col = -1
            for j in range(num_vars):
                if art[i][j] != 0 and all(b != j for b in art_basis):
                    new_col = j
                    break
            if new_col == -1:
                return ARTIFICIAL_STUCK, (it1, fix, 0)
            n_pivot_row(art, art_basis, new_col, i)
            fix += 1
    for r in range(m):                                             # :437-441
        for c in range(num_vars):
            main[r][c] = art[r][c]
        main[r][num_vars] = art[r][num_art_vars]
    for i in range(m):                                             # :444-451
        bc = art_basis[i]
        main_basis[i] = bc
        scale = main[m][bc]
        if scale != 0:
            for c in range(num_vars + 1):
                main[m][c] -= scale * main[i][c]
    st, it2 = n_solve_tableau(main, main_basis, is_max, rule)      # :452
    return st, (it1, fix, it2)
