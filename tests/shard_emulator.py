"""TEST-ONLY stand-in for `_ffi.DeviceTableau(shard=...)` that runs the sharded per-iteration
protocol of libb200lp.so (k_enter on the objective replica, local k_ratio, k_cand, all-gather
of candidates, k_winner, local k_pivot) on the CPU with oracle primitives and a
torch.distributed (gloo) all-gather.  Lets the N > 1 host logic and the exchange protocol be
checked without GPUs: the pivot trace must equal the unsharded oracle's."""
import numpy as np
import torch.distributed as dist

from oracle import oracle


class _Res:
    def __init__(self, iterations):
        self.iterations = iterations


class OracleShard:
    def __init__(self, R, C, is_max=True, opts=None, shard=None):
        self.R, self.C, self.is_max = R, C, is_max
        self.rank, self.world, self.uid = shard
        assert len(self.uid) == 128
        self.tol = 1024.0
        self.trace = []
        per = -(-(R - 1) // self.world)
        self.row0 = min(R - 1, per * self.rank)

    def upload(self, blk, basis):
        self.blk, self.basis = blk.copy(), basis.copy()

    def iterate(self, max_iters=0):
        blk, basis, C = self.blk, self.basis, self.C
        ml = blk.shape[0] - 1
        thr = (self.tol / 2.0) * oracle.cl_epsilon()
        it = 0
        while True:
            j = oracle.find_entering_column(blk, self.is_max, self.tol)       # replica: no exchange
            if j < 0:
                return 0, _Res(it), self.trace
            col = blk[:, j].copy()                                           # k_ratio snapshot
            best = (None, None, -1, None)                                    # (q, key, row, cand)
            for i in range(ml):
                if col[i] > thr:
                    q = blk[i, C - 1] / col[i]
                    if best[2] < 0 or q < best[0]:
                        best = (q, self.row0 + i, self.row0 + i, None)
            if best[2] >= 0:
                src = blk[best[2] - self.row0]
                best = best[:3] + (src / src[j],)                            # k_cand
            gathered = [None] * self.world
            dist.all_gather_object(gathered, best)                            # the one exchange
            win = None
            for g in gathered:                                               # k_winner
                if g[2] >= 0 and (win is None or g[0] < win[0] or (g[0] == win[0] and g[1] < win[1])):
                    win = g
            if win is None:
                return 1, _Res(it), self.trace
            p, prow = win[2], win[3]
            self.trace.append((j, p))
            prod = col[:, None] * prow[None, :]                              # rounded product ...
            new = blk - prod                                                 # ... then rounded difference
            pl = p - self.row0
            if 0 <= pl < ml:
                new[pl] = prow
                basis[pl] = j
            blk[:] = new
            it += 1

    def download_solution(self):
        return self.blk[:, self.C - 1].copy(), self.blk[-1].copy(), self.basis.copy()

    def close(self):
        pass
