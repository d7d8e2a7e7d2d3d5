"""TEST-ONLY stand-in for `_ffi.DeviceTableau(shard=...)` that runs the sharded per-iteration
protocol of libb200lp.so's k_iter on the CPU, with numpy for the arithmetic and torch.distributed
(gloo) for the exchange, so the N > 1 logic can be checked without GPUs:

  * the LOOKAHEAD: iteration k+1 is decided (entering column on the objective replica, local
    ratio test, scaled candidate row) from the block as it was BEFORE pivot k plus pivot k's
    (column snapshot, scaled row) -- a - t*p evaluated lazily must equal the updated cell;
  * the two-step exchange: every rank's (ratio, key, row) header to every rank, all pick the same
    lexicographic minimum, then ONLY the winner's scaled row travels;
  * the update applied afterwards to the whole block.

The pivot trace must equal the unsharded oracle's."""
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

    def _lazy(self, cells, rows, cols, pending):
        """Cells blk[rows, cols] as they will be after the pending pivot, without applying it."""
        if pending is None:
            return cells
        col, prow, pl = pending
        out = cells - col[rows] * prow[cols]                  # rounded product, rounded difference
        if np.ndim(rows) == 0:
            return prow[cols] if rows == pl else out
        if np.ndim(cols) == 0:
            out = out.copy()
            if 0 <= pl < len(out):
                out[pl] = prow[cols]
        return out

    def iterate(self, max_iters=0):
        blk, basis, C = self.blk, self.basis, self.C
        ml = blk.shape[0] - 1
        rows_all = np.arange(ml + 1)
        thr_e = (self.tol / 8.0) * oracle.cl_epsilon()
        thr_p = (self.tol / 2.0) * oracle.cl_epsilon()
        it = 0
        pending = None                                            # (col, prow, p_local) of pivot k
        while True:
            # ---- look(k -> k+1) on the not-yet-updated block ------------------------------------
            obj = self._lazy(blk[ml, :C - 1], ml, np.arange(C - 1), pending)
            key = obj if self.is_max else -obj
            j = int(np.argmin(key))                                   # first minimum
            entering = key[j] < 0.0 - thr_e
            if entering:
                col = self._lazy(blk[:, j], rows_all, j, pending)     # snapshot incl. objective row
                rhs = self._lazy(blk[:, C - 1], rows_all, C - 1, pending)
                hdr = (None, None, -1)
                for i in range(ml):
                    if col[i] > 0.0 + thr_p:
                        q = rhs[i] / col[i]
                        if hdr[2] < 0 or q < hdr[0]:
                            hdr = (q, self.row0 + i, self.row0 + i)
                headers = [None] * self.world
                dist.all_gather_object(headers, hdr)                  # step 1: headers everywhere
                win, winner = None, -1
                for g, h in enumerate(headers):
                    if h[2] >= 0 and (win is None or h[0] < win[0] or (h[0] == win[0] and h[1] < win[1])):
                        win, winner = h, g
                box = [None]
                if win is not None:
                    if winner == self.rank:                           # step 2: only the winner's row
                        i = win[2] - self.row0
                        row = self._lazy(blk[i], i, np.arange(C), pending)
                        box = [row / col[i]]
                    dist.broadcast_object_list(box, src=winner)
            # ---- update(k): retire the pending pivot ---------------------------------------------
            if pending is not None:
                pcol, prow, pl = pending
                new = blk - pcol[:, None] * prow[None, :]
                if 0 <= pl < ml:
                    new[pl] = prow
                blk[:] = new
                it += 1
            if not entering:
                return 0, _Res(it), self.trace
            if win is None:
                return 1, _Res(it), self.trace
            p = win[2]
            self.trace.append((j, p))
            pl = p - self.row0
            if 0 <= pl < ml:
                basis[pl] = j
            pending = (col, box[0], pl if 0 <= pl < ml else -1)

    def download_solution(self):
        return self.blk[:, self.C - 1].copy(), self.blk[-1].copy(), self.basis.copy()

    def close(self):
        pass
