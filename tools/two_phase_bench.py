"""One timing line for the two-phase path (src/simplex.lisp:402-452) at BASELINE scale (dev tool):
a mixed <= / >= / = LP with m constraint rows through b200lp_solve_two_phase on 1..N GPUs of this
process; the N-GPU result must equal the 1-GPU result bit for bit.

    python tools/two_phase_bench.py [m n] [ndev ...]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import _ffi  # noqa: E402


def mixed_lp(m, n, seed=1234):
    """A ~ U[0,1), a feasible point x0 ~ U[0,1); one third each of <= (slack), >= (surplus +
    artificial) and = (artificial) rows built around x0; max c.x with c ~ U[0,1) -- bounded because
    every column has positive entries in the <= rows."""
    rng = np.random.default_rng(seed)
    A = rng.random((m, n))
    x0 = rng.random(n)
    kinds = rng.integers(0, 3, size=m)
    ax = A @ x0
    rhs = np.where(kinds == 0, ax + rng.uniform(n / 16.0, n / 8.0, m),
                   np.where(kinds == 1, ax * rng.uniform(0.5, 1.0, m), ax))
    c = rng.random(n)
    n_slack = int((kinds != 2).sum())
    art_rows = np.flatnonzero(kinds != 0)
    C = n + n_slack + 1
    main = np.zeros((m + 1, C))
    main[:m, :n] = A
    main[:m, -1] = rhs
    mb = np.full(m, C, np.int32)
    off = np.cumsum(kinds != 2) - 1
    le, ge = np.flatnonzero(kinds == 0), np.flatnonzero(kinds == 1)
    main[le, n + off[le]] = 1.0
    mb[le] = n + off[le]
    main[ge, n + off[ge]] = -1.0
    main[m, :n] = -c
    na = len(art_rows)
    art = np.zeros((m + 1, C + na))
    art[:m, :C - 1] = main[:m, :C - 1]
    art[:m, -1] = main[:m, -1]
    ab = mb.copy()
    rev = art_rows[::-1]                       # the reference pushes rows (:258, 295-299)
    art[rev, C - 1 + np.arange(na)] = 1.0
    ab[rev] = C - 1 + np.arange(na)
    art[m, :C - 1] = art[art_rows][:, :C - 1].sum(axis=0)
    art[m, -1] = art[art_rows][:, -1].sum()
    return art, ab, main, mb


def main():
    args = [int(a) for a in sys.argv[1:]]
    m, n = (args[0], args[1]) if len(args) >= 2 else (8192, 8192)
    ndevs = args[2:] or [1]
    art0, ab0, main0, mb0 = mixed_lp(m, n)
    base = None
    for nd in ndevs:
        art, ab, mn, mb = art0.copy(), ab0.copy(), main0.copy(), mb0.copy()
        t0 = time.perf_counter()
        st, res = _ffi.solve_two_phase(art, ab, mn, mb, True, _ffi.make_opts(devices=list(range(nd))))
        wall = time.perf_counter() - t0
        piv = int(res.iterations_phase1 + res.iterations_cleanup + res.iterations)
        same = None
        if base is None:
            base = (mn[:, -1].copy(), mn[-1].copy(), mb.copy())
        else:
            same = bool(np.array_equal(mn[:, -1], base[0]) and np.array_equal(mn[-1], base[1])
                        and np.array_equal(mb, base[2]))
        print(json.dumps(dict(workload=f"mixed <=/>=/= LP m={m} n={n}, two-phase", n_devices=nd, status=int(st),
                              art_shape=list(art.shape), main_shape=list(mn.shape),
                              pivots_phase1=int(res.iterations_phase1), pivots_cleanup=int(res.iterations_cleanup),
                              pivots_phase2=int(res.iterations), objective=float(res.objective),
                              wall_s=wall, ms_device_loops=float(res.ms_solve),
                              pivots_per_s=piv / wall, equals_one_gpu_result=same)), flush=True)


if __name__ == "__main__":
    main()
