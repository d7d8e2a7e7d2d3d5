// small.cuh -- a whole solve in ONE launch of ONE CTA, the tableau resident in shared memory.
//
// The reference's usual diet is a handful of variables (README.md:30-62; every branch-and-bound
// node, src/simplex.lisp:506-542).  For a tableau that fits the 227 KB of shared memory of one SM
// (up to about 28 000 cells: 3 x 6 ... 100 x 270) the per-pivot launches, the grid barriers and the
// ping-pong HBM buffers of the big path are all overhead: k_small loads the tableau once (zero-copy
// from the caller's pinned staging buffer), runs n-solve-tableau (src/simplex.lisp:399-461) --
// single tableau or the two-phase list branch including the artificial clean-up and the
// objective re-pricing -- with block-level barriers only, and writes back what the accessors read.
// One kernel launch and one stream synchronisation per b200lp_solve / b200lp_solve_two_phase call.
//
// Same arithmetic contract as everywhere (__ddiv_rn, __dmul_rn then __dsub_rn, strict compares,
// lowest index wins): bit-identical to the big path and to the oracle.
#pragma once
#include "kernels.cuh"

namespace b200lp {

constexpr int kSmallThreads = 512;
constexpr int ST_INFEASIBLE = 2;             // == B200LP_INFEASIBLE
constexpr int ST_ART_STUCK = 4;              // == B200LP_ARTIFICIAL_STUCK
constexpr int ST_ART_NONZERO = 5;            // == B200LP_ARTIFICIAL_NONZERO

struct SmallHeader {                         // written to the (mapped, pinned) output buffer
    int status;
    int wrote;                               // 1: out_rhs / out_obj / out_basis (and out_full) are valid
    long long iters;                         // single tableau: pivots; two-phase: phase-2 pivots
    long long iters_phase1, iters_cleanup, redundant;
    double objective;
};

struct SmallArgs {
    // input (host-mapped): tableau R x C_in packed (ld == C_in), basis[R-1]; two-phase: the
    // artificial tableau is the input tableau and main_obj[C] is the main objective row
    const double *in_tab;
    const int32_t *in_basis;
    const double *main_obj;                  // two-phase only
    int R, C_in, C_main;                     // C_main == C_in for a single tableau
    int two_phase, is_max, rule, feas_reference;
    double thr_enter, thr_pivot, thr_feas;   // thr_feas: tol * eps (unscaled)
    long long max_iters;                     // per phase; 0 = unlimited
    // output (host-mapped)
    SmallHeader *hdr;
    double *out_rhs;                         // R
    double *out_obj;                         // C_main
    int32_t *out_basis;                      // R-1
    double *out_full;                        // R x C_main or null: the whole solved (main) tableau
    double *out_art_full;                    // two-phase + writeback: solved phase-1 tableau, R x C_in
    int32_t *out_art_basis;
    int2 *trace;                             // single tableau only
    int trace_cap;
};

// Shared-memory footprint of k_small for a given shape (bytes); the host uses the same formula.
__host__ __device__ inline size_t small_smem_bytes(int R, int C_in, int C_main)
{
    const size_t ldS = (size_t)(C_in | 1);
    size_t b = (size_t)R * ldS * 8;          // tableau
    b += (size_t)R * 8;                      // pivot-column snapshot
    b += (size_t)C_main * 8;                 // main objective row (two-phase)
    b += (size_t)R * 4;                      // basis
    b += (size_t)C_in;                       // is-basic flags
    return (b + 15) & ~(size_t)15;
}

template <int THREADS>
__device__ __forceinline__ Cand small_block_min(Cand c, Cand *red)
{
    return cand_block_min<THREADS>(c, red);
}

struct SmallState {
    double *T;
    double *colS;
    int *basis;
    int ldS, m;
};

// n-pivot-row (src/simplex.lisp:337-359) on the shared-memory tableau, width C.
__device__ __forceinline__ void small_pivot(const SmallState &S, const int C, const int j, const int p)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSmallThreads / 32;
    // snapshot of column j over all rows (each a[r,j] is read before row r changes, :353)
    for (int r = tid; r <= S.m; r += kSmallThreads) S.colS[r] = S.T[r * S.ldS + j];
    __syncthreads();
    const double s = S.colS[p];
    double *prow = S.T + p * S.ldS;
    for (int c = tid; c < C; c += kSmallThreads) prow[c] = __ddiv_rn(prow[c], s);
    __syncthreads();
    for (int r = warp; r <= S.m; r += NW) {
        if (r == p) continue;
        const double t = S.colS[r];
        double *row = S.T + r * S.ldS;
        for (int c = lane; c < C; c += 32) row[c] = __dsub_rn(row[c], __dmul_rn(t, prow[c]));
    }
    if (tid == 0) S.basis[p] = j;
    __syncthreads();
}

// n-solve-tableau, single-tableau branch (src/simplex.lisp:453-461) on width C (nv = C - 1).
// Returns the status; *iters_out = pivots done.
__device__ __forceinline__ int small_solve(const SmallState &S, const int C, const int is_max,
                                           const int rule, const double thr_enter,
                                           const double thr_pivot, const long long max_iters,
                                           Cand *red, int2 *trace, const int trace_cap,
                                           long long *iters_out)
{
    const int tid = threadIdx.x;
    const int nv = C - 1, rhs = C - 1;
    long long it = 0;
    int status = ST_OPTIMAL;
    for (;;) {
        // find-entering-column :362-379
        Cand best;
        best.q = 0.0; best.key = 0; best.row = -1;
        const double *obj = S.T + S.m * S.ldS;
        for (int c = tid; c < nv; c += kSmallThreads) {
            const double v = obj[c];
            const double k = is_max ? v : -v;
            if (rule == 0) {
                if (best.row < 0 || k < best.q) { best.q = k; best.key = c; best.row = c; }
            } else if (best.row < 0 && k < 0.0 - thr_enter) {
                best.q = 0.0; best.key = c; best.row = c;
            }
        }
        best = small_block_min<kSmallThreads>(best, red);
        __syncthreads();                                           // red[] reuse below
        const bool accept = (best.row >= 0) && (rule != 0 || best.q < 0.0 - thr_enter);
        if (!accept) break;
        if (max_iters > 0 && it >= max_iters) { status = ST_ITERATION_LIMIT; break; }
        const int j = best.row;
        // find-pivoting-row :382-389
        Cand c;
        c.q = 0.0; c.key = 0; c.row = -1;
        for (int i = tid; i < S.m; i += kSmallThreads) {
            const double a = S.T[i * S.ldS + j];
            if (0.0 + thr_pivot < a) {
                Cand d;
                d.q = __ddiv_rn(S.T[i * S.ldS + rhs], a);
                d.key = rule ? S.basis[i] : i;
                d.row = i;
                c = cand_min(c, d);
            }
        }
        c = small_block_min<kSmallThreads>(c, red);
        __syncthreads();
        if (c.row < 0) { status = ST_UNBOUNDED; break; }
        if (tid == 0 && trace && it < trace_cap) trace[it] = make_int2(j, c.row);
        small_pivot(S, C, j, c.row);
        ++it;
    }
    *iters_out = it;
    return status;
}

__global__ void __launch_bounds__(kSmallThreads, 1) k_small(const SmallArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Cand red[kSmallThreads / 32];
    const int tid = threadIdx.x;
    const int R = A.R, m = R - 1;
    SmallState S;
    S.ldS = A.C_in | 1;
    S.m = m;
    S.T = reinterpret_cast<double *>(smem_raw);
    S.colS = S.T + (size_t)R * S.ldS;
    double *mobj = S.colS + R;
    S.basis = reinterpret_cast<int *>(mobj + A.C_main);
    unsigned char *isb = reinterpret_cast<unsigned char *>(S.basis + R);

    // ---- load (zero-copy from the pinned staging buffer) -----------------------------------------
    for (int k = tid; k < R * A.C_in; k += kSmallThreads) {
        const int r = k / A.C_in, c = k - r * A.C_in;
        S.T[r * S.ldS + c] = A.in_tab[k];
    }
    for (int i = tid; i < m; i += kSmallThreads) S.basis[i] = A.in_basis[i];
    if (A.two_phase)
        for (int c = tid; c < A.C_main; c += kSmallThreads) mobj[c] = A.main_obj[c];
    __syncthreads();

    SmallHeader h;
    h.status = ST_OPTIMAL; h.wrote = 0; h.iters = 0; h.iters_phase1 = 0; h.iters_cleanup = 0;
    h.redundant = 0; h.objective = 0.0;
    int C = A.C_in;                                                // width of the tableau being solved
    if (A.two_phase) {
        // ---- phase 1: the artificial tableau is a `min` problem (:317-319, 402-404) -------------
        const int nv = A.C_main - 1, art_nv = A.C_in - 1;
        const double obj0 = S.T[m * S.ldS + art_nv];
        const double thr_feas = A.feas_reference ? A.thr_feas : A.thr_feas * fmax(1.0, fabs(obj0));
        long long it1 = 0;
        int st = small_solve(S, A.C_in, /*is_max=*/0, A.rule, A.thr_enter, A.thr_pivot, A.max_iters,
                             red, nullptr, 0, &it1);
        h.iters_phase1 = it1;
        if (st == ST_OPTIMAL) {
            // (unless (fp= 0 obj tol) (error 'infeasible-problem-error)) :405-407
            const double obj = S.T[m * S.ldS + art_nv];
            if (!(fabs(0.0 - obj) <= thr_feas)) st = ST_INFEASIBLE;
        }
        // ---- zero-level artificials out of the basis :419-434 ------------------------------------
        for (int i = 0; i < m && st == ST_OPTIMAL; ++i) {
            if (S.basis[i] < nv) continue;                         // uniform: shared memory
            const double rhs = S.T[i * S.ldS + art_nv];
            if (A.feas_reference ? (rhs != 0.0) : !(fabs(rhs) <= thr_feas)) { st = ST_ART_NONZERO; break; }
            for (int c = tid; c < A.C_in; c += kSmallThreads) isb[c] = 0;
            __syncthreads();
            for (int k = tid; k < m; k += kSmallThreads) isb[S.basis[k]] = 1;
            __syncthreads();
            Cand best;
            best.q = 0.0; best.key = 0; best.row = -1;
            const double *row = S.T + i * S.ldS;
            for (int c = tid; c < nv; c += kSmallThreads) {
                if (isb[c]) continue;
                const double a = row[c];
                Cand d;
                d.key = c; d.row = c;
                if (A.feas_reference) {
                    if (a == 0.0) continue;
                    d.q = 0.0;                                     // first non-zero: lowest index
                } else {
                    if (!(fabs(a) > A.thr_pivot)) continue;
                    d.q = -fabs(a);                                // largest magnitude, first index on ties
                }
                best = cand_min(best, d);
            }
            best = small_block_min<kSmallThreads>(best, red);
            __syncthreads();
            if (best.row < 0) {
                if (A.feas_reference) { st = ST_ART_STUCK; break; }
                h.redundant += 1;                                  // a combination of the other rows
                continue;
            }
            small_pivot(S, A.C_in, best.row, i);
            h.iters_cleanup += 1;
        }
        if (st == ST_OPTIMAL && A.out_art_full) {                  // the solved phase-1 tableau, on request
            for (int k = tid; k < R * A.C_in; k += kSmallThreads) {
                const int r = k / A.C_in, c = k - r * A.C_in;
                A.out_art_full[k] = S.T[r * S.ldS + c];
            }
            for (int i = tid; i < m; i += kSmallThreads) A.out_art_basis[i] = S.basis[i];
        }
        if (st == ST_OPTIMAL) {
            // ---- coefficients + RHS into the main tableau (:437-441): same rows, RHS column moves
            __syncthreads();
            for (int r = tid; r < m; r += kSmallThreads) S.T[r * S.ldS + nv] = S.T[r * S.ldS + art_nv];
            for (int c = tid; c <= nv; c += kSmallThreads) S.T[m * S.ldS + c] = mobj[c];
            __syncthreads();
            // ---- re-price the objective row, rows in order (:444-451) ---------------------------
            double *obj = S.T + m * S.ldS;
            for (int i = 0; i < m; ++i) {
                const int bc = S.basis[i];
                if (bc >= nv) continue;                            // redundant row: nothing to price
                const double scale = obj[bc];
                __syncthreads();                                   // everyone has read obj[bc]
                if (scale != 0.0) {
                    const double *row = S.T + i * S.ldS;
                    for (int c = tid; c <= nv; c += kSmallThreads)
                        obj[c] = __dsub_rn(obj[c], __dmul_rn(scale, row[c]));
                }
                __syncthreads();
            }
            C = A.C_main;
            long long it2 = 0;
            st = small_solve(S, C, A.is_max, A.rule, A.thr_enter, A.thr_pivot, A.max_iters, red,
                             nullptr, 0, &it2);
            h.iters = it2;
        }
        h.status = st;
        if (st != ST_OPTIMAL && st != ST_UNBOUNDED && st != ST_ITERATION_LIMIT) C = 0;   // nothing to write back
        else if (C != A.C_main) C = 0;                             // phase 1 ended the solve
    } else {
        long long it = 0;
        h.status = small_solve(S, C, A.is_max, A.rule, A.thr_enter, A.thr_pivot, A.max_iters, red,
                               A.trace, A.trace_cap, &it);
        h.iters = it;
    }

    // ---- write back what the accessors read (src/simplex.lisp:74-120) --------------------------------
    if (C > 0) {
        h.wrote = 1;
        h.objective = S.T[m * S.ldS + (C - 1)];
        for (int r = tid; r < R; r += kSmallThreads) A.out_rhs[r] = S.T[r * S.ldS + (C - 1)];
        for (int c = tid; c < C; c += kSmallThreads) A.out_obj[c] = S.T[m * S.ldS + c];
        for (int i = tid; i < m; i += kSmallThreads) A.out_basis[i] = S.basis[i];
        if (A.out_full)
            for (int k = tid; k < R * C; k += kSmallThreads) {
                const int r = k / C, c = k - r * C;
                A.out_full[k] = S.T[r * S.ldS + c];
            }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        *A.hdr = h;
    }
}

} // namespace b200lp
