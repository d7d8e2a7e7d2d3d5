"""A/B of the two loop packagings on a resident tableau (dev tool, run under gpurun).

    python tools/loop_ab.py [--shapes cfg3,slab8,cfg2,small] [--variants 10,20,21,22] [--look 0,4,8,16]
                            [--iters 400] [--tag NAME]

variant 1..13  = one k_iter launch per pivot (kernels.cuh), tile shape per solver.cu's table
variant 20..23 = k_persist, one cooperative kernel for the whole call (persist.cuh)
--look N       = B200LP_LOOK_CTAS for the persistent loop (0 = library default)
Prints one JSON line per (shape, variant, look): ms per iteration, algorithmic GB/s = 16*R*C/t,
pivots/s and, for the persistent loop, the decision chain's breakdown in us per pivot.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import _ffi, synthetic  # noqa: E402

SHAPES = {
    "cfg3": (8192, 16384),      # BASELINE config 3
    "slab8": (1024, 23552),     # one rank's share of config 3 on 8 GPUs (same row length)
    "slab8c4": (2048, 47104),   # one rank's share of config 4 on 8 GPUs
    "cfg2": (1024, 2048),       # BASELINE config 2 (L2 resident)
    "cfg5": (4096, 4096),
    "small": (256, 512),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="cfg3,slab8,cfg2,small")
    ap.add_argument("--variants", default="10,20")
    ap.add_argument("--look", default="0")
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--tag", default="loop_ab")
    args = ap.parse_args()
    rows = []
    for shape in args.shapes.split(","):
        m, n = SHAPES[shape]
        tab, basis = synthetic.dense_tableau(m, n)
        R, C = tab.shape
        for v in [int(x) for x in args.variants.split(",")]:
            looks = [int(x) for x in args.look.split(",")] if v >= 20 else [0]
            for g in looks:
                if g:
                    os.environ["B200LP_LOOK_CTAS"] = str(g)
                else:
                    os.environ.pop("B200LP_LOOK_CTAS", None)
                opts = _ffi.make_opts(pivot_variant=v)
                with _ffi.DeviceTableau(R, C, True, opts) as d:
                    d.upload(tab, basis)
                    d.iterate(10)
                    t0 = time.perf_counter()
                    st, res, _ = d.iterate(args.iters)
                    wall = time.perf_counter() - t0
                    it = max(int(res.iterations), 1)
                    n_look = max(int(res.look_kernel_launches), 1)
                    row = dict(shape=shape, m=m, n=n, variant=v, loop_mode=int(res.loop_mode),
                               look_ctas=int(res.look_ctas), look_cluster=int(res.look_cluster), status=int(st), pivots=it,
                               us_iter=1e3 * res.ms_solve / it,
                               gbs=16.0 * R * C * it / res.ms_solve / 1e6,
                               pivots_per_s=1e3 * it / res.ms_solve, wall_ms=1e3 * wall,
                               launches=int(res.kernel_launches),
                               us_look=1e3 * res.ms_look_kernel / n_look)
                    if res.loop_mode == 2:
                        row.update(us_look_wait=1e3 * res.ms_look_wait / n_look,
                                   us_look_ratio=1e3 * res.ms_look_ratio / n_look,
                                   us_look_push=1e3 * res.ms_look_push / n_look,
                                   us_look_peer_wait=1e3 * res.ms_look_peer_wait / n_look,
                                   us_look_row=1e3 * res.ms_look_row / n_look,
                                   sm_mhz=res.sm_clock_mhz,
                                   us_dbg=[round(1e3 * x / n_look, 2) for x in res.ms_look_dbg[:5]])
                    rows.append(row)
                    print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/{args.tag}.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
