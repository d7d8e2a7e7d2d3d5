"""Full solve of one BASELINE config on the GPU through b200lp_solve, with a summary line (dev tool).

usage: python tools/solve_config.py cfg5 [rule] [max_iters] [ndev] [zero_frac]   (ndev > 1: one process, several GPUs)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import _ffi, synthetic  # noqa: E402

CONFIGS = {"cfg2": (1024, 2048, False), "cfg3": (8192, 16384, False),
           "cfg4": (16384, 32768, False), "cfg5": (4096, 4096, True)}


def main():
    name = sys.argv[1]
    rule = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ndev = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    m, n, deg = CONFIGS[name]
    zf = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
    tab, basis = synthetic.dense_tableau(m, n, degenerate=deg, zero_frac=zf)
    t0 = time.perf_counter()
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(pivot_rule=rule, max_iters=cap,
                                                          devices=list(range(ndev))))
    wall = time.perf_counter() - t0
    print(json.dumps(dict(config=name, m=m, n=n, rule=rule, zero_frac=zf, devices_in_one_process=ndev,
                          exchange_mode=int(res.exchange_mode), status=st, status_text=_ffi.strerror(st),
                          pivots=int(res.iterations), objective=res.objective, wall_s=wall,
                          pivots_per_s=res.iterations / wall, ms_h2d=res.ms_h2d,
                          ms_solve=res.ms_solve, ms_d2h=res.ms_d2h,
                          gbs=16.0 * (m + 1) * (n + m + 1) * res.iterations / res.ms_solve / 1e6)),
          flush=True)


if __name__ == "__main__":
    main()
