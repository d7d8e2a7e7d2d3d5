"""Config 5 family with a FIXED number of cone rows (16 / 64) at m = 256, 512, 1024 under both rules, with
the oracle's cycle probe (dev tool; TEST INFRASTRUCTURE -- uses oracle/).  Output: profiles/r02_cfg5_cycle_probe_64_cone_rows.jsonl"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from linear_programming_b200 import synthetic
from oracle import oracle
oracle.set_num_threads(8)
for m in (256, 512, 1024):
    for cone in (16, 64):
        for rule in (0, 1):
            tab,basis = synthetic.dense_tableau(m, m, degenerate=True, zero_frac=cone/m)
            t0=time.time()
            st, info = oracle.solve_cycle_probe(tab, basis, True, rule=rule, max_iters=400000, parallel=m>=512)
            print(json.dumps(dict(m=m, cone=cone, rule=rule, st=st, obj=float(tab[-1,-1]), s=round(time.time()-t0,1), **info)), flush=True)
