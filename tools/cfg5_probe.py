"""Config 5 on the CPU oracle with the cycle probe: for the degenerate generator, does the reference
rule cycle, stall or finish, and does Bland finish?  (dev tool; TEST INFRASTRUCTURE -- uses oracle/)

    python tools/cfg5_probe.py 256,512 0.5,0.125,0.03125 [cap]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linear_programming_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256").split(",")]
fracs = [float(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0.5").split(",")]
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
oracle.build()
oracle.set_num_threads(len(os.sched_getaffinity(0)))
for m in sizes:
    for zf in fracs:
        for rule in (0, 1):
            tab, basis = synthetic.dense_tableau(m, m, degenerate=True, zero_frac=zf)
            t0 = time.time()
            st, info = oracle.solve_cycle_probe(tab, basis, True, rule=rule, max_iters=cap, parallel=m >= 512)
            print(json.dumps(dict(m=m, zero_frac=zf, rule=rule, status=st, objective=float(tab[-1, -1]),
                                  seconds=round(time.time() - t0, 2), **info)), flush=True)
