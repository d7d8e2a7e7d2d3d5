"""Run the CPU oracle (oracle/simplex_oracle.c, the restatement of src/simplex.lisp:337-461) to
the END of a BASELINE config and commit what the solution accessors read as a golden fixture.

    python tools/make_full_goldens.py cfg3            -> tests/golden/cfg3_final.npz
    python tools/make_full_goldens.py cfg5 --rule 1   -> tests/golden/cfg5_rule1_final.npz

TEST INFRASTRUCTURE: this is the generating script of tests/golden/*_final.npz (the GPU box has
no /root/reference and no time budget for an 18 751-pivot CPU solve; the fixture travels instead).
The fixture holds: final basis (int32[m]), RHS column (f64[R]), objective row (f64[C]), the full
pivot trace (int32[iters, 2] = entering column, leaving row), pivot count, status, and the
generator arguments.  Everything is compared bit for bit by tests/test_gpu_parity.py and by
bench.py's `parity` block.
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402

CONFIGS = {
    "cfg2": dict(m=1024, n=2048, degenerate=False),
    "cfg3": dict(m=8192, n=16384, degenerate=False),
    "cfg5": dict(m=4096, n=4096, degenerate=True),
}


def trace_digest(trace):
    """sha256 over the int32 (j, r) pairs -- the same digest bench.py prints for the GPU run."""
    return hashlib.sha256(np.ascontiguousarray(trace, dtype=np.int32).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--rule", type=int, default=0)
    ap.add_argument("--max-iters", type=int, default=0)
    ap.add_argument("--zero-frac", type=float, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    kw = {}
    if args.zero_frac is not None:
        kw["zero_frac"] = args.zero_frac
    tab, basis = synthetic.dense_tableau(cfg["m"], cfg["n"], degenerate=cfg["degenerate"], **kw)
    R, C = tab.shape
    oracle.build()
    threads = oracle.set_num_threads(len(os.sched_getaffinity(0)))
    cap = 4_000_000
    t0 = time.time()
    st, iters, trace = oracle.solve(tab, basis, True, rule=args.rule, max_iters=args.max_iters,
                                    parallel=True, trace_cap=cap)
    dt = time.time() - t0
    assert iters <= cap
    trace = np.asarray(trace, dtype=np.int32).reshape(-1, 2)
    name = args.out or os.path.join(
        ROOT, "tests", "golden",
        f"{args.config}{'_rule1' if args.rule else ''}_final.npz")
    np.savez_compressed(
        name, basis=basis, rhs=np.ascontiguousarray(tab[:, C - 1]), obj_row=np.ascontiguousarray(tab[R - 1]),
        trace=trace, iterations=np.int64(iters), status=np.int32(st),
        objective=np.float64(tab[R - 1, C - 1]), m=np.int64(cfg["m"]), n=np.int64(cfg["n"]),
        seed=np.int64(1234), rule=np.int32(args.rule), degenerate=np.bool_(cfg["degenerate"]),
        zero_frac=np.float64(args.zero_frac if args.zero_frac is not None else 0.5),
        trace_sha256=np.str_(trace_digest(trace)))
    print(f"{args.config} rule {args.rule}: status {st}, {iters} pivots, objective "
          f"{tab[R - 1, C - 1]!r}, {dt:.1f} s on {threads} threads -> {name} "
          f"({os.path.getsize(name)} bytes)")


if __name__ == "__main__":
    main()
