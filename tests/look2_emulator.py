"""TEST-ONLY numpy model of the round-2 look role (persist.cuh: k_iter2 / k_persist), so that the
algebra the GPU relies on is also checked on the CPU, sharded or not:

  * the look role never reads the objective row or the RHS column from the tableau again after
    the first pivot: it keeps COMPACT running copies, updated with the same rounded product and
    rounded difference the update tiles apply to the tableau -- they must stay bit-identical;
  * pivot k is decided from S_{k-2} (the tableau two pivots back) plus pivot k-1's column
    snapshot and scaled row, evaluated lazily;
  * phase A: column gather + ratio test -> this rank's candidate; every rank then scales ITS OWN
    candidate row speculatively and all of them are exchanged at once (one exchange per pivot);
    every rank picks the same lexicographic (ratio, key) minimum;
  * phase B: the winner's scaled row updates the compact objective row, whose argmin is the next
    entering column.

Ranks are run one after the other in this process (the protocol is deterministic); the pivot
trace, the final tableau and the basis must equal the unsharded oracle's bit for bit."""
import numpy as np

from oracle import oracle


def _partition(m, world, rank):
    per = -(-m // max(world, 1))
    b = min(m, per * rank)
    return b, min(m, b + per)


def solve(tab, basis, is_max=True, rule=0, max_iters=0, world=1, tol=1024.0):
    """Returns (status, trace, final tableau, final basis)."""
    R, C = tab.shape
    m, nv, rhs_col = R - 1, C - 1, C - 1
    thr_e = (tol / 8.0) * oracle.cl_epsilon()
    thr_p = (tol / 2.0) * oracle.cl_epsilon()
    ranks = []
    for g in range(world):
        b, e = _partition(m, world, g)
        blk = np.vstack([tab[b:e], tab[m:m + 1]]).copy()           # rows + objective replica
        ranks.append(dict(b=b, e=e, S=[blk, blk.copy()], basis=basis[b:e].copy(),
                          objc=blk[-1].copy(), rhsc=blk[:, rhs_col].copy()))
    sign = 1.0 if is_max else -1.0

    def entering(objc):
        key = sign * objc[:nv]
        if rule == 0:
            j = int(np.argmin(key))                                # first minimum
            return j if key[j] < 0.0 - thr_e else -1
        hits = np.flatnonzero(key < 0.0 - thr_e)
        return int(hits[0]) if hits.size else -1

    j = entering(ranks[0]["objc"])
    trace, prev = [], None                                         # prev = pivot k-1: (j, p, cols, prow)
    status, k = 0, 1
    while True:
        if j < 0:
            status = 0
            break
        if max_iters and len(trace) >= max_iters:
            status = 3
            break
        cands, cols = [], []
        for g, rk in enumerate(ranks):
            src = rk["S"][k & 1] if k >= 2 else rk["S"][0]          # S_{k-2} (S_0 for k = 1)
            ml = rk["e"] - rk["b"]
            a = src[:, j].copy()
            bvec = rk["rhsc"]
            if prev is not None:
                pj, pb = prev[3][j], prev[3][rhs_col]
                t = prev[2][g]
                a = a - t * pj
                bvec = bvec - t * pb
                pl = prev[1] - rk["b"]
                if 0 <= pl < ml:
                    a[pl], bvec[pl] = pj, pb
                rk["rhsc"] = bvec                                  # RHS column of S_{k-1}
            cols.append(a)
            best = None
            for i in range(ml):
                if 0.0 + thr_p < a[i]:
                    q = bvec[i] / a[i]
                    key = rk["b"] + i
                    if rule:
                        key = prev[0] if (prev is not None and prev[1] == rk["b"] + i) else int(rk["basis"][i])
                    if best is None or q < best[0] or (q == best[0] and key < best[1]):
                        best = (q, key, rk["b"] + i)
            # speculative: this rank's own candidate row, scaled
            row = None
            if best is not None:
                i = best[2] - rk["b"]
                row = src[i].copy()
                if prev is not None:
                    row = prev[3].copy() if prev[1] == best[2] else row - prev[2][g][i] * prev[3]
                row = row / a[i]
            cands.append((best, row))
        win = None
        for g, (best, row) in enumerate(cands):                    # same choice on every rank
            if best is not None and (win is None or best[0] < win[0][0] or
                                     (best[0] == win[0][0] and best[1] < win[0][1])):
                win = (best, row)
        if win is None:
            status = 1
            break
        p, prow = win[0][2], win[1]
        trace.append((j, p))
        for g, rk in enumerate(ranks):                             # phase B on every replica
            tobj = cols[g][-1]
            rk["objc"] = rk["objc"] - tobj * prow
            pl = p - rk["b"]
            if 0 <= pl < rk["e"] - rk["b"]:
                rk["basis"][pl] = j
        jn = entering(ranks[0]["objc"])
        assert all(entering(rk["objc"]) == jn for rk in ranks)
        # the update tiles of pivot k-1 (concurrent with this decision on the GPU): S_{k-2} -> S_{k-1}
        if prev is not None:
            for g, rk in enumerate(ranks):
                _apply(rk, k - 1, prev, g)
        prev = (j, p, cols, prow)
        j = jn
        k += 1
    if prev is not None:                                           # the last decided pivot
        for g, rk in enumerate(ranks):
            _apply(rk, k - 1, prev, g)
    last = (k - 1) & 1 if prev is not None else 0
    out = np.vstack([rk["S"][last][:-1] for rk in ranks] + [ranks[0]["S"][last][-1:]])
    for rk in ranks:                                               # the compact copies ARE the tableau's
        assert np.array_equal(rk["objc"], rk["S"][last][-1]), "compact objective row diverged"
    return status, trace, out, np.concatenate([rk["basis"] for rk in ranks])


def _apply(rk, kk, piv, g):
    """update(kk): S_{kk-1} -> S_{kk} (n-pivot-row part 2, src/simplex.lisp:349-358)."""
    src = rk["S"][(kk - 1) & 1] if kk >= 2 else rk["S"][0]
    col, prow, p = piv[2][g], piv[3], piv[1]
    dst = src - col[:, None] * prow[None, :]
    pl = p - rk["b"]
    if 0 <= pl < rk["e"] - rk["b"]:
        dst[pl] = prow
    rk["S"][kk & 1] = dst
