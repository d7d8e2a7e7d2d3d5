"""Wall time of one-shot b200lp_solve calls on tiny problems (dev tool): where the fixed cost goes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import _ffi, synthetic  # noqa: E402

for m, n in [(2, 3), (64, 96), (256, 512)]:
    tab0, basis0 = synthetic.dense_tableau(m, n, seed=3)
    _ffi.solve(tab0.copy(), basis0.copy(), True)
    t0 = time.perf_counter()
    N = 50
    for _ in range(N):
        st, res, _ = _ffi.solve(tab0.copy(), basis0.copy(), True)
    dt = (time.perf_counter() - t0) / N
    print(f"m={m} n={n}: {1e3 * dt:.3f} ms per solve call, {res.iterations} pivots, "
          f"in-call {res.ms_total:.3f} ms (h2d {res.ms_h2d:.3f}, loop {res.ms_solve:.3f}, d2h {res.ms_d2h:.3f})",
          flush=True)
