"""Synthetic dense LPs of BASELINE.json's configs: max c.x s.t. Ax <= b, x >= 0.

numpy default_rng(seed) (PCG64), draws in the order A (row-major m x n), b, c; all fp64.
The tableau is what the reference's build-tableau produces for such a problem
(src/simplex.lisp:214-287): rows 0..m-1 = [A | I_m | b], row m = [-c | 0 | 0], basis[i] = n+i.
"""
import numpy as np


def dense_lp(m, n, seed=1234, degenerate=False, zero_frac=0.5):
    """A ~ U[0,1), b ~ U[n/8, 3n/8), c ~ U[0,1).

    degenerate=True is config 5: small-integer data with exact ratio ties.  The first half of
    the rows are cone constraints through the origin (A in {-1,0,1}, b = 0: every ratio there
    is exactly 0, so the leaving row is decided by the tie-break alone); the second half
    (A in {0,1,2}, b in {n/4, n/2, 3n/4}) bounds the polytope; c in {1,2,3}.  `zero_frac` is the
    share of cone rows.  Measured with the oracle's cycle probe (tools/cfg5_probe.py; DESIGN.md
    section 2): neither rule ever revisits a basis on this family -- the reference rule STALLS, it
    does not cycle -- and the length of the stall grows steeply with the number of cone rows:
    zero_frac 1/2 at m = 4096 is still at objective 0 after 3 * 10^6 pivots under either rule
    (prefix-parity use only), while zero_frac 1/64 (64 cone rows, the instance bench.py and the
    full-size parity fixture call config 5) solves in 53 616 pivots under the reference rule.
    Bland's rule is slow on the whole family (about m^2.6 pivots: 177 000 at m = 1024)."""
    rng = np.random.default_rng(seed)
    if degenerate:
        A = rng.integers(0, 3, size=(m, n)).astype(np.float64)
        b = rng.integers(1, 4, size=m).astype(np.float64) * n / 4.0
        c = rng.integers(1, 4, size=n).astype(np.float64)
        half = int(m * zero_frac)
        A[:half] = rng.integers(-1, 2, size=(half, n))
        b[:half] = 0.0
        return A, b, c
    A = rng.random((m, n))
    b = rng.uniform(n / 8.0, 3.0 * n / 8.0, m)
    c = rng.random(n)
    return A, b, c


def tableau_from_lp(A, b, c, out=None):
    """[A | I | b ; -c | 0 | 0] as a row-major (m+1) x (n+m+1) fp64 matrix plus the slack basis."""
    m, n = A.shape
    R, C = m + 1, n + m + 1
    tab = out if out is not None else np.zeros((R, C))
    if out is not None:
        tab[:] = 0.0
    tab[:m, :n] = A
    tab[np.arange(m), n + np.arange(m)] = 1.0
    tab[:m, C - 1] = b
    tab[m, :n] = -c
    basis = np.arange(n, n + m, dtype=np.int32)
    return tab, basis


def dense_tableau(m, n, seed=1234, degenerate=False, out=None, zero_frac=0.5):
    A, b, c = dense_lp(m, n, seed, degenerate, zero_frac)
    return tableau_from_lp(A, b, c, out=out)


def dense_block(m, n, row_begin, row_end, seed=1234, out=None):
    """Rows [row_begin, row_end) of dense_tableau(m, n, seed) plus its objective row, WITHOUT
    materialising the other rows: the (row_end - row_begin + 1) x (n + m + 1) block a rank of a
    row-block sharded solve uploads, and its slice of the slack basis.  PCG64 is advanced past
    the rows that are skipped (one 64-bit draw per double), so the values are bit-identical to
    the full generator's (tests/test_host_frontend.py)."""
    rng = np.random.default_rng(seed)
    rows = row_end - row_begin
    C = n + m + 1
    blk = out if out is not None else np.zeros((rows + 1, C))
    if out is not None:
        blk[:] = 0.0
    rng.bit_generator.advance(row_begin * n)
    step = max(1, (1 << 24) // max(n, 1))                  # ~128 MB of draws at a time
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        blk[r0:r1, :n] = rng.random((r1 - r0, n))
    rng.bit_generator.advance((m - row_end) * n)
    b = rng.uniform(n / 8.0, 3.0 * n / 8.0, m)
    c = rng.random(n)
    blk[np.arange(rows), n + row_begin + np.arange(rows)] = 1.0
    blk[:rows, C - 1] = b[row_begin:row_end]
    blk[rows, :n] = -c
    basis = np.arange(n + row_begin, n + row_end, dtype=np.int32)
    return blk, basis


README_LP = dict(A=np.array([[2.0, 1.0, 0.0], [0.0, 1.0, 1.0]]), b=np.array([8.0, 7.0]),
                 c=np.array([1.0, 4.0, 3.0]))  # README.md:30-62 -> obj 57/2, x=(1/2, 7, 0)
