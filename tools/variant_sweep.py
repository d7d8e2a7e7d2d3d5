"""Times the rank-1 pivot kernel variants on a resident tableau (dev tool, run under gpurun).

usage: python tools/variant_sweep.py [m n] [variants...]
Prints per variant: avg k_pivot time (CUDA events around each launch), achieved algorithmic
GB/s = 16*R*C / t, whole-iteration time and pivots/s.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import _ffi, synthetic  # noqa: E402


def main():
    args = [int(a) for a in sys.argv[1:]]
    m, n = (args[0], args[1]) if len(args) >= 2 else (8192, 16384)
    variants = args[2:] or [1, 2, 3, 4, 5, 6, 7, 8, 9]
    t0 = time.time()
    tab, basis = synthetic.dense_tableau(m, n)
    print(f"# LP m={m} n={n} built in {time.time() - t0:.1f}s, tableau {tab.nbytes / 1e9:.3f} GB", flush=True)
    R, C = tab.shape
    rows = []
    for v in variants:
        opts = _ffi.make_opts(time_kernels=os.environ.get('SWEEP_TIME_KERNELS', '1') != '0', pivot_variant=v)
        with _ffi.DeviceTableau(R, C, True, opts) as d:
            t0 = time.time()
            d.upload(tab, basis)
            up = time.time() - t0
            d.iterate(5)
            st, res, _ = d.iterate(40)
            per = res.ms_pivot_kernel / max(res.pivot_kernel_launches, 1) or res.ms_solve / res.iterations
            row = dict(variant=v, pivots=res.iterations, ms_iter=res.ms_solve / res.iterations,
                       ms_pivot=per, gbs=16.0 * R * C / per / 1e6,
                       pivots_per_s=1e3 * res.iterations / res.ms_solve, upload_s=round(up, 3),
                       us_look=1e3 * res.ms_look_kernel / max(res.look_kernel_launches, 1),
                       launches=res.kernel_launches)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/variant_sweep_{m}x{n}.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
