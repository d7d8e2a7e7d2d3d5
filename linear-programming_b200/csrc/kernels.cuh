// kernels.cuh -- sm_100a kernels of the dense tableau simplex iteration.
//
// The solve loop (b200lp_iterate / b200lp_solve*) runs ONE kernel per iteration, k_iter
// (second half of this file): lookahead CTAs + rank-1 update tiles, with the sharded candidate
// exchange done inside the kernel over peer-mapped buffers; k_look + k_update are the same two
// roles as separate kernels for the NCCL fallback.
//
// The step-by-step entry points (b200lp_find_entering_column, _find_pivoting_row, _pivot) map
// one reference function (src/simplex.lisp) to one small kernel each (first half of this file):
//   k_enter   find-entering-column  :362-379   reduced-cost row scan, (value,index) argmin
//   k_ratio   find-pivoting-row     :382-389   strided column gather + ratio argmin; also
//                                              snapshots the pivot column a[:,j] (n-pivot-row
//                                              reads each a[r,j] before row r changes, :353)
//   k_cand    n-pivot-row part 1    :344-348   candidate pivot row / pivot element
//   k_winner  (sharded only)                   pick the global leaving row from all ranks
//   k_pivot   n-pivot-row part 2    :349-358   rank-1 update in place
//
// Arithmetic contract (bit-identical to the reference's double-float path and to
// oracle/simplex_oracle.c): __ddiv_rn for the row scale and the ratios, __dmul_rn then
// __dsub_rn for the update (never contracted into an FMA), strict compares with lowest-index
// tie-breaks.  Every kernel returns at once when the device status word is not RUNNING, so
// the host can enqueue iterations ahead of knowing whether the solve has finished.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lp {

constexpr int ST_RUNNING = 100;          // device-only; public codes are in b200lp.h
constexpr int ST_OPTIMAL = 0;
constexpr int ST_UNBOUNDED = 1;
constexpr int ST_ITERATION_LIMIT = 3;

constexpr int kEnterThreads = 1024;
constexpr int kRatioThreads = 256;
constexpr int kPivotThreads = 256;
constexpr int kCandHdr = 4;              // doubles in front of a candidate row (32 B)

// Device-resident loop state; one per shard.
struct alignas(16) DevState {
    int status;               // ST_RUNNING or a final code
    int j;                    // entering column of the current iteration
    int p;                    // leaving row (GLOBAL row index) of the current iteration
    int winner;               // rank whose candidate row won (0 when unsharded)
    long long iters;          // pivots completed
    long long max_iters;      // 0 = unlimited
    double q;                 // winning ratio
    unsigned int ticket;      // last-block ticket of k_ratio
    int forced_row;           // >= 0: b200lp_pivot / clean-up pivots skip the ratio argmin
};

// Header in front of each candidate pivot row (exchanged between ranks when sharded).
struct alignas(16) CandHdr {
    double q;                 // b_r / a[r,j]
    long long key;            // tie-break key: global row (reference rule) or basis[r] (Bland)
    long long row;            // global row index, -1 = this rank has no eligible row
    long long pad;
};
static_assert(sizeof(CandHdr) == kCandHdr * sizeof(double), "header size");

struct Cand { double q; int key; int row; };

// (q, key) lexicographic minimum; row < 0 marks "no candidate".  Reproduces the sequential
// `finding i minimizing` rule: strict <, first index wins (key == global row for rule 0).
__device__ __forceinline__ Cand cand_min(const Cand a, const Cand b)
{
    if (b.row < 0) return a;
    if (a.row < 0) return b;
    if (b.q < a.q || (b.q == a.q && b.key < a.key)) return b;
    return a;
}

__device__ __forceinline__ Cand cand_shfl_xor(const Cand c, int lane_mask)
{
    Cand o;
    o.q = __shfl_xor_sync(0xffffffffu, c.q, lane_mask);
    o.key = __shfl_xor_sync(0xffffffffu, c.key, lane_mask);
    o.row = __shfl_xor_sync(0xffffffffu, c.row, lane_mask);
    return o;
}

__device__ __forceinline__ Cand cand_warp_min(Cand c)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c = cand_min(c, cand_shfl_xor(c, s));
    return c;
}

template <int THREADS>
__device__ __forceinline__ Cand cand_block_min(Cand c, Cand *smem /* THREADS/32 */)
{
    c = cand_warp_min(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = c;
    __syncthreads();
    if (warp == 0) {
        Cand d;
        d.q = 0.0; d.key = 0; d.row = -1;
        if (lane < THREADS / 32) d = smem[lane];
        d = cand_warp_min(d);
        if (lane == 0) smem[0] = d;
    }
    __syncthreads();
    return smem[0];
}

// ------------------------------------------------------------------------------------------
// k_enter: find-entering-column (src/simplex.lisp:362-379).  One CTA scans the objective row.
// max problem: first argmin, accept iff value < -(tol/8)eps; min problem: first argmax, accept
// iff value > +(tol/8)eps -- folded into one argmin over key = is_max ? v : -v (negation is
// exact and order reversing).  Bland: lowest index whose key passes the same threshold.
// Also enforces the iteration cap (build extension) once an entering column exists.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEnterThreads)
k_enter(const double *__restrict__ obj, int nv, int is_max, double thr, int rule, DevState *st)
{
    if (st->status != ST_RUNNING) return;
    __shared__ Cand red[kEnterThreads / 32];
    Cand best;
    best.q = 0.0; best.key = 0; best.row = -1;
    if (rule == 0) {
        for (int i = threadIdx.x; i < nv; i += kEnterThreads) {
            const double v = obj[i];
            const double k = is_max ? v : -v;
            if (best.row < 0 || k < best.q) { best.q = k; best.key = i; best.row = i; }
        }
    } else {
        for (int i = threadIdx.x; i < nv; i += kEnterThreads) {
            const double v = obj[i];
            const double k = is_max ? v : -v;
            if (k < 0.0 - thr) { best.q = 0.0; best.key = i; best.row = i; break; }
        }
    }
    best = cand_block_min<kEnterThreads>(best, red);
    if (threadIdx.x == 0) {
        const bool accept = (best.row >= 0) && (rule != 0 || best.q < 0.0 - thr);
        if (!accept) {
            st->status = ST_OPTIMAL;
            st->j = -1;
        } else if (st->max_iters > 0 && st->iters >= st->max_iters) {
            st->status = ST_ITERATION_LIMIT;
            st->j = best.row;
        } else {
            st->j = best.row;
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_ratio: find-pivoting-row (src/simplex.lisp:382-389) over this shard's rows, fused with the
// pivot-column snapshot colbuf[i] = a[i,j] for every local row including the objective replica.
// Eligible rows: a[i,j] > (tol/2)eps; ratio = rhs_i / a[i,j] (true division, no clamp).
// Multi-CTA; the last CTA to finish (ticket) reduces the per-CTA partials and publishes the
// shard's candidate.  Unsharded (world == 1) it also finalises the iteration: leaving row,
// UNBOUNDED status, pivot trace.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRatioThreads)
k_ratio(const double *__restrict__ tab, int64_t ld, int m_local, int R_local, int rhs_col,
        const int32_t *__restrict__ basis, int row0, double thr, int rule, int world,
        double *__restrict__ colbuf, DevState *st, Cand *__restrict__ partials,
        CandHdr *__restrict__ hdr, int2 *__restrict__ trace, int trace_cap)
{
    if (st->status != ST_RUNNING) return;
    __shared__ Cand red[kRatioThreads / 32];
    __shared__ bool is_last;
    const int j = st->j;
    const int forced = st->forced_row;
    const int i = blockIdx.x * kRatioThreads + threadIdx.x;
    Cand c;
    c.q = 0.0; c.key = 0; c.row = -1;
    if (i < R_local) {
        const double a = tab[(int64_t)i * ld + j];
        colbuf[i] = a;
        if (i < m_local) {
            if (forced >= 0) {
                if (row0 + i == forced) { c.q = 0.0; c.key = 0; c.row = forced; }
            } else if (0.0 + thr < a) {
                c.q = __ddiv_rn(tab[(int64_t)i * ld + rhs_col], a);
                c.key = rule ? basis[i] : row0 + i;
                c.row = row0 + i;
            }
        }
    }
    c = cand_block_min<kRatioThreads>(c, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = c;
        __threadfence();
        const unsigned int t = atomicAdd(&st->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Cand d;
    d.q = 0.0; d.key = 0; d.row = -1;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kRatioThreads) {
        // volatile: written by other CTAs of this launch
        const volatile Cand *vp = partials + b;
        Cand o;
        o.q = vp->q; o.key = vp->key; o.row = vp->row;
        d = cand_min(d, o);
    }
    __syncthreads();
    d = cand_block_min<kRatioThreads>(d, red);
    if (threadIdx.x == 0) {
        st->ticket = 0;
        hdr->q = d.q; hdr->key = d.key; hdr->row = d.row; hdr->pad = 0;
        if (world == 1) {
            if (d.row < 0) {
                st->status = ST_UNBOUNDED;
                st->p = -1;
            } else {
                st->p = d.row;
                st->q = d.q;
                st->winner = 0;
                if (trace && st->iters < trace_cap) trace[st->iters] = make_int2(j, d.row);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_cand: n-pivot-row part 1 (src/simplex.lisp:344-348): this shard's candidate row divided by
// its pivot element, written behind the header; pad columns [C, ld) are zeroed so the update
// may run over whole 16-byte vectors.  A shard without an eligible row writes nothing.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cand(const double *__restrict__ tab, int64_t ld, int C, int row0, const DevState *st,
       const CandHdr *__restrict__ hdr, double *__restrict__ cand_row)
{
    if (st->status != ST_RUNNING) return;
    const long long row = hdr->row;
    if (row < 0) return;
    const double *src = tab + (row - row0) * ld;
    const double s = src[st->j];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < (int)ld; c += gridDim.x * blockDim.x)
        cand_row[c] = (c < C) ? __ddiv_rn(src[c], s) : 0.0;
}

// ------------------------------------------------------------------------------------------
// k_winner (sharded only): after the all-gather every rank holds all candidates; each picks the
// same (ratio, key) lexicographic minimum, i.e. the row the unsharded scan would have picked.
// ------------------------------------------------------------------------------------------
__global__ void k_winner(const double *__restrict__ gathered, int64_t stride, int world,
                         DevState *st, int2 *__restrict__ trace, int trace_cap)
{
    if (st->status != ST_RUNNING) return;
    if (threadIdx.x != 0) return;
    Cand best;
    best.q = 0.0; best.key = 0; best.row = -1;
    int w = -1;
    for (int g = 0; g < world; ++g) {
        const CandHdr *h = reinterpret_cast<const CandHdr *>(gathered + g * stride);
        Cand c;
        c.q = h->q; c.key = (int)h->key; c.row = (int)h->row;
        const Cand nb = cand_min(best, c);
        if (nb.row != best.row) { best = nb; w = g; }
    }
    if (best.row < 0) {
        st->status = ST_UNBOUNDED;
        st->p = -1;
    } else {
        st->p = best.row;
        st->q = best.q;
        st->winner = w;
        if (trace && st->iters < trace_cap) trace[st->iters] = make_int2(st->j, best.row);
    }
}

// ------------------------------------------------------------------------------------------
// k_pivot: n-pivot-row part 2 (src/simplex.lisp:349-358), the bandwidth-bound hot kernel.
//   row p            <- scaled pivot row
//   every other row r <- a[r,:] - colbuf[r] * prow[:]   (objective replica included)
// In place; each element is read once and written once: 16*R*C algorithmic bytes.
// Tiling: CTA = TR rows x (256 threads * VEC double2) columns; every thread keeps its slice of
// the pivot row in registers and streams UNROLL rows of 16-byte loads before the stores.
// ------------------------------------------------------------------------------------------
template <bool STREAM>
__device__ __forceinline__ double2 ld_tab(const double2 *p)
{
    if (STREAM) return __ldcs(p);
    return *p;
}
template <bool STREAM>
__device__ __forceinline__ void st_tab(double2 *p, const double2 v)
{
    if (STREAM) __stcs(p, v);
    else *p = v;
}

template <int TR, int UNROLL, int VEC, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads)
k_pivot(double *__restrict__ tab, int64_t ld, int m_local, int R_local, int row0,
        const double *__restrict__ colbuf, const double *__restrict__ cand_base, int64_t cand_stride,
        int32_t *__restrict__ basis, DevState *st)
{
    if (st->status != ST_RUNNING) return;
    __shared__ double s_col[TR];
    const int ldv = (int)(ld >> 1);                       // row length in double2
    const int r_begin = blockIdx.y * TR;
    const int r_end = min(R_local, r_begin + TR);
    for (int t = threadIdx.x; t < TR; t += kPivotThreads)
        s_col[t] = (r_begin + t < r_end) ? colbuf[r_begin + t] : 0.0;
    const int p_global = st->p;
    const int p_rel = p_global - row0;
    const int p_local = (p_rel >= 0 && p_rel < m_local) ? p_rel : -1;
    const double2 *prow2 =
        reinterpret_cast<const double2 *>(cand_base + (int64_t)st->winner * cand_stride + kCandHdr);
    int cv[VEC];
    double2 pr[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        cv[v] = (blockIdx.x * VEC + v) * kPivotThreads + threadIdx.x;
        pr[v] = (cv[v] < ldv) ? prow2[cv[v]] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    double2 *tab2 = reinterpret_cast<double2 *>(tab);
    for (int r = r_begin; r < r_end; r += UNROLL) {
        double2 a[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (r + u < r_end && cv[v] < ldv)
                    a[u][v] = ld_tab<STREAM>(tab2 + (int64_t)(r + u) * ldv + cv[v]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (r + u < r_end) {
                const double t = s_col[r + u - r_begin];
                const bool is_p = (r + u == p_local);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (cv[v] < ldv) {
                        double2 o;
                        o.x = __dsub_rn(a[u][v].x, __dmul_rn(t, pr[v].x));
                        o.y = __dsub_rn(a[u][v].y, __dmul_rn(t, pr[v].y));
                        if (is_p) o = pr[v];
                        st_tab<STREAM>(tab2 + (int64_t)(r + u) * ldv + cv[v], o);
                    }
                }
            }
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        if (p_local >= 0) basis[p_local] = st->j;         // (setf (aref basis p) j) :358
        st->iters += 1;
    }
}

// ==========================================================================================
// The pipelined loop (b200lp_iterate / b200lp_solve*).
//
// The decision chain of an iteration (entering column -> ratio test -> scaled pivot row) needs
// only O(R + C) cells of the tableau, and every one of them can be computed from the tableau
// BEFORE the previous pivot's rank-1 update plus that pivot's (column, scaled row):
//     a'[r,c] = a[r,c] - col[r] * prow[c]        (r != p)         a'[p,c] = prow[c]
// -- the same rounded product and rounded difference the update writes, so the bits agree.  Hence
//     look(k -> k+1)  decides iteration k+1 from the tableau S_{k-1} and pivot k's (col, prow)
//     update(k)       streams S_{k-1} -> S_k (out of place, ping-pong buffers)
// have the same inputs and run CONCURRENTLY: the latency-bound chain -- and, when the tableau is
// row-block sharded, the exchange of candidate pivot rows between GPUs -- hides behind the
// HBM-bound update.  Two packagings of the same two device functions:
//   k_iter            ONE persistent kernel per iteration: the first few CTAs play the look role,
//                     the rest pull update tiles from a counter.  Sharded, the look role exchanges
//                     candidates itself through peer-mapped buffers over NVLink (header from
//                     every rank, then the winner's scaled row to every rank) -- no NCCL call,
//                     no second stream, nothing on the host between iterations.
//   k_look + k_update the same roles as two kernels on two streams with an NCCL all-gather in
//                     between (used when peer mapping is unavailable).
// Per-iteration decisions live in a 4-slot ring (slot = k & 3): IterState, pivot column,
// candidate headers/row.
// ==========================================================================================
constexpr int ST_START = 101;            // ring slot holds no pending pivot (first look of a call)
constexpr int ST_PEER_TIMEOUT = -7;      // == B200LP_ERR_PEER_TIMEOUT
constexpr int ST_BARRIER_TIMEOUT = -6;   // == B200LP_ERR_INTERNAL: a look-grid barrier gave up
constexpr int kRing = 4;
constexpr int kLookThreads = 256;        // == kPivotThreads: both roles share k_iter's CTA shape
constexpr int kLookMaxCtas = 32;
constexpr int kLookBatch = 16;           // row cells a look thread loads before using the first
constexpr int kLookRows = 8;             // tableau rows a look thread gathers per batch
constexpr int kMaxWorld = 8;

struct alignas(16) IterState {
    int status;               // ST_RUNNING: pivot (j, p) is pending
    int j;                    // entering column
    int p;                    // leaving row (global), -1 = this shard's candidate only (NCCL path)
    int w;                    // rank whose scaled row is the pivot row
    long long iters;          // pivots completed before this iteration
    long long pad;
};

struct alignas(16) Report {   // what the host polls
    int status;
    int pad;
    long long iters;
};

// Scratch of one look grid: arrival counters of its grid-wide barriers (re-armed by the last CTA
// to leave), the per-CTA partial results of the two argmin stages, the update-tile counters.
struct alignas(16) LookSync {
    unsigned int bar[4];
    unsigned int fail;
    unsigned int look_count;              // telemetry: look roles that ran to a decision
    unsigned long long look_t0;           //            %globaltimer at the start of the current one
    unsigned long long look_ns;           //            summed duration
    Cand part_enter[kLookMaxCtas];
    Cand part_ratio[kLookMaxCtas];
};

// Peer-mapped exchange buffer, one per rank, identical layout everywhere (8-byte words):
//   hdr  [kRing][kMaxWorld] CandHdr   candidate header of every rank
//   row  [kRing][ld]        double    the winner's scaled pivot row
//   flag1[kRing][kMaxWorld] u64       "header of rank r for this slot has landed" (sequence no.)
//   flag2[kRing]            u64       "row for this slot has landed"
struct Xchg {
    double *peer[kMaxWorld];  // base of every rank's buffer as mapped into THIS process
    int rank;
    int world;
    int64_t ld;
    unsigned long long epoch; // high bits of the sequence numbers of this b200lp_iterate call
};
__host__ __device__ __forceinline__ int64_t xchg_hdr_off(int slot, int r) { return (slot * kMaxWorld + r) * kCandHdr; }
__host__ __device__ __forceinline__ int64_t xchg_row_off(int slot, int64_t ld) { return kRing * kMaxWorld * kCandHdr + slot * ld; }
__host__ __device__ __forceinline__ int64_t xchg_flag1_off(int slot, int r, int64_t ld) { return kRing * kMaxWorld * kCandHdr + kRing * ld + slot * kMaxWorld + r; }
__host__ __device__ __forceinline__ int64_t xchg_flag2_off(int slot, int64_t ld) { return kRing * kMaxWorld * kCandHdr + kRing * ld + kRing * kMaxWorld + slot; }
__host__ __device__ __forceinline__ int64_t xchg_words(int64_t ld) { return kRing * kMaxWorld * kCandHdr + kRing * ld + kRing * kMaxWorld + kRing + 4; }

// Pick the global leaving row from candidate headers (hdr_stride doubles apart).
// Same (ratio, key) lexicographic minimum on every rank == the unsharded first-index scan.
__device__ __forceinline__ Cand resolve_winner(const double *hdr_base, int64_t hdr_stride, int world,
                                               int *winner)
{
    Cand best;
    best.q = 0.0; best.key = 0; best.row = -1;
    int w = 0;
    for (int g = 0; g < world; ++g) {
        const double *h = hdr_base + g * hdr_stride;
        Cand c;
        c.q = __ldcg(h);
        c.key = (int)__ldcg(reinterpret_cast<const long long *>(h) + 1);
        c.row = (int)__ldcg(reinterpret_cast<const long long *>(h) + 2);
        const Cand nb = cand_min(best, c);
        if (nb.row != best.row) { best = nb; w = g; }
    }
    *winner = w;
    return best;
}

struct LookArgs {
    const double *src;        // tableau S_{k-1}: this shard's block, R_local x ld
    int64_t ld;
    int C, m_local, R_local, row0, world, rank;
    int is_max, rule;
    int slot_in, slot_out;    // ring slots of iteration k (pending) and k+1 (decided here)
    double thr_enter, thr_pivot;
    long long max_iters;      // 0 = unlimited
    IterState *ring;          // kRing slots
    double *colring;          // kRing x col_stride: pivot columns
    int64_t col_stride;
    // candidates: exchange mode 0 (one shard) / 1 (NCCL all-gather after the kernel):
    double *candring;         // kRing x cand_stride: this shard's (hdr + scaled row)
    const double *gathring;   // mode 1: kRing x world x cand_stride, filled by the all-gather
    int64_t cand_stride;      // kCandHdr + ld
    // exchange mode 2 (peer-mapped, in-kernel): xchg.peer[] != nullptr
    Xchg xchg;
    int mode;
    int32_t *basis;
    Report *report;
    int2 *trace;
    int trace_cap;
    LookSync *sync;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Grid-wide barrier of the G look CTAs (co-resident by construction: the lowest block indices
// of k_iter, or a small grid on a high-priority stream).  Data written before it is visible
// after it to loads that bypass L1.
// The spin is bounded (SM cycle counter, 2 cycles per ns of `timeout_ns`): should the CTAs ever
// not be co-resident (MPS, a debugger, a scheduler that dispatches differently) the wait gives up,
// raises bit 1 of *fail and the call ends with B200LP_ERR_INTERNAL instead of hanging the GPU.
__device__ __forceinline__ void look_barrier(unsigned int *ctr, int G, unsigned int *fail,
                                             unsigned long long timeout_ns)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        const long long budget = (long long)(timeout_ns << 1);
        while (*reinterpret_cast<volatile unsigned int *>(ctr) < (unsigned)G) {
            if (*reinterpret_cast<volatile unsigned int *>(fail) & 2u) break;
            if (clock64() - t0 > budget) { atomicOr(fail, 2u); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

// Every CTA reduces the same per-CTA partials in the same order: identical result everywhere.
__device__ __forceinline__ Cand look_reduce_partials(const Cand *part, int G, Cand *s_out)
{
    if (threadIdx.x < 32) {
        Cand d;
        d.q = 0.0; d.key = 0; d.row = -1;
        if ((int)threadIdx.x < G) {
            const Cand *p = part + threadIdx.x;
            d.q = __ldcg(&p->q); d.key = __ldcg(&p->key); d.row = __ldcg(&p->row);
        }
        d = cand_warp_min(d);
        if (threadIdx.x == 0) *s_out = d;
    }
    __syncthreads();
    return *s_out;
}

// Spin until *flag == want (a peer's system-scope store); false on timeout.
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long want,
                                          unsigned long long timeout_ns)
{
    const volatile unsigned long long *f = flag;
    if (*f == want) return true;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        for (int k = 0; k < 64; ++k)
            if (*f == want) return true;
        if (global_timer_ns() - t0 > timeout_ns) return false;
    }
}

// ------------------------------------------------------------------------------------------
// look role: find-entering-column + find-pivoting-row + the division half of n-pivot-row
// (src/simplex.lisp:362-389, 344-348) for iteration k+1, evaluated on the not-yet-updated
// tableau.  It also retires iteration k: basis[p] <- j (:358), pivot trace, iteration count.
// G CTAs x 256 threads split each scan; three grid barriers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void look_role(const LookArgs &A, const int cta, const int G)
{
    __shared__ Cand red[kLookThreads / 32];
    __shared__ Cand s_part;
    __shared__ int s_p, s_w, s_ok;
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const IterState st = A.ring[A.slot_in];
    IterState *st_out = A.ring + A.slot_out;
    if (cta == 0 && tid == 0) A.sync->look_t0 = global_timer_ns();
    if (st.status != ST_RUNNING && st.status != ST_START) {   // solve already over: pass it on
        if (cta == 0 && tid == 0) *st_out = st;
        return;
    }
    const bool pending = st.status == ST_RUNNING;
    long long iters = st.iters;
    int p_local = -1;
    const double *prow = nullptr;
    if (pending) {
        int p = st.p, w = st.w;
        if (A.mode == 1) {                                     // verdict of the all-gather
            if (tid == 0) {
                int ww = 0;
                s_p = resolve_winner(A.gathring + (int64_t)A.slot_in * A.world * A.cand_stride,
                                     A.cand_stride, A.world, &ww).row;
                s_w = ww;
            }
            __syncthreads();
            p = s_p; w = s_w;
            if (p < 0) {                                       // no rank had an eligible row
                if (cta == 0 && tid == 0) {
                    IterState o = st;
                    o.status = ST_UNBOUNDED;
                    *st_out = o;
                    A.report->iters = iters;
                    A.report->status = ST_UNBOUNDED;
                }
                return;
            }
            prow = A.gathring + ((int64_t)A.slot_in * A.world + w) * A.cand_stride + kCandHdr;
        } else if (A.mode == 2) {
            prow = A.xchg.peer[A.rank] + xchg_row_off(A.slot_in, A.ld);
        } else {
            prow = A.candring + (int64_t)A.slot_in * A.cand_stride + kCandHdr;
        }
        const int rel = p - A.row0;
        if (rel >= 0 && rel < A.m_local) p_local = rel;
        if (cta == 0 && tid == 0) {
            if (p_local >= 0) A.basis[p_local] = st.j;         // (setf (aref basis p) j) :358
            if (A.trace && iters < A.trace_cap) A.trace[iters] = make_int2(st.j, p);
        }
        iters += 1;
    }
    const double *col = A.colring + (int64_t)A.slot_in * A.col_stride;
    double *col_out = A.colring + (int64_t)A.slot_out * A.col_stride;
    double *cand_out = A.candring + (int64_t)A.slot_out * A.cand_stride;
    const int nv = A.C - 1;
    int fin = ST_RUNNING;
    int j = -1, win_rank = 0;
    Cand c;
    c.q = 0.0; c.key = 0; c.row = -1;

    // ---- stage 1: entering column over the (updated) objective row ---------------------------
    {
        const double *obj = A.src + (int64_t)A.m_local * A.ld;
        const double t_obj = pending ? col[A.m_local] : 0.0;
        const int chunk = (nv + G - 1) / G;
        const int c_end = min(nv, (cta + 1) * chunk);
        Cand best;
        best.q = 0.0; best.key = 0; best.row = -1;
        // kLookBatch loads of a thread are in flight together: the scan costs one memory round
        // trip per batch instead of one per element (this chain is pure latency)
        for (int base = cta * chunk + tid; base < c_end; base += kLookBatch * kLookThreads) {
            double vo[kLookBatch], vp[kLookBatch];
#pragma unroll
            for (int u = 0; u < kLookBatch; ++u) {
                const int cc = base + u * kLookThreads;
                vo[u] = cc < c_end ? obj[cc] : 0.0;
                vp[u] = (pending && cc < c_end) ? prow[cc] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kLookBatch; ++u) {
                const int cc = base + u * kLookThreads;
                if (cc >= c_end) break;
                double v = vo[u];
                if (pending) v = __dsub_rn(v, __dmul_rn(t_obj, vp[u]));
                const double k = A.is_max ? v : -v;
                if (A.rule == 0) {
                    if (best.row < 0 || k < best.q) { best.q = k; best.key = cc; best.row = cc; }
                } else if (best.row < 0 && k < 0.0 - A.thr_enter) {
                    best.q = 0.0; best.key = cc; best.row = cc;   // Bland: lowest passing index
                }
            }
        }
        best = cand_block_min<kLookThreads>(best, red);
        if (G > 1) {
            if (tid == 0) A.sync->part_enter[cta] = best;
            look_barrier(&A.sync->bar[0], G, &A.sync->fail, A.timeout_ns);
            best = look_reduce_partials(A.sync->part_enter, G, &s_part);
        }
        const bool accept = (best.row >= 0) && (A.rule != 0 || best.q < 0.0 - A.thr_enter);
        if (!accept) fin = ST_OPTIMAL;
        else if (A.max_iters > 0 && iters >= A.max_iters) fin = ST_ITERATION_LIMIT;
        j = accept ? best.row : -1;
    }

    if (fin == ST_RUNNING) {
        // ---- stage 2: pivot-column snapshot + ratio test over this shard's rows ----------------
        const int rhs = A.C - 1;
        const double pj = pending ? prow[j] : 0.0;
        const double pb = pending ? prow[rhs] : 0.0;
        const int chunk = (A.R_local + G - 1) / G;
        const int r_end = min(A.R_local, (cta + 1) * chunk);
        for (int base = cta * chunk + tid; base < r_end; base += kLookRows * kLookThreads) {
            double va[kLookRows], vb[kLookRows], vt[kLookRows];
#pragma unroll
            for (int u = 0; u < kLookRows; ++u) {
                const int i = base + u * kLookThreads;
                const bool in = i < r_end;
                const double *rowp = A.src + (int64_t)(in ? i : 0) * A.ld;
                va[u] = in ? rowp[j] : 0.0;
                vb[u] = in ? rowp[rhs] : 0.0;
                vt[u] = (pending && in) ? col[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kLookRows; ++u) {
                const int i = base + u * kLookThreads;
                if (i >= r_end) break;
                double a = va[u], b = vb[u];
                if (pending) {
                    const double t = vt[u];
                    a = (i == p_local) ? pj : __dsub_rn(a, __dmul_rn(t, pj));
                    b = (i == p_local) ? pb : __dsub_rn(b, __dmul_rn(t, pb));
                }
                col_out[i] = a;
                if (i < A.m_local && 0.0 + A.thr_pivot < a) {
                    Cand d;
                    d.q = __ddiv_rn(b, a);
                    d.key = A.rule ? ((i == p_local) ? st.j : A.basis[i]) : A.row0 + i;
                    d.row = A.row0 + i;
                    c = cand_min(c, d);
                }
            }
        }
        __syncthreads();                                       // red[] reuse
        c = cand_block_min<kLookThreads>(c, red);
        if (G > 1) {
            if (tid == 0) A.sync->part_ratio[cta] = c;
            look_barrier(&A.sync->bar[1], G, &A.sync->fail, A.timeout_ns);
            c = look_reduce_partials(A.sync->part_ratio, G, &s_part);
        }
        // c = this shard's best (ratio, key, row)

        bool i_win = true;                                     // modes 0/1: always scale own row
        if (A.mode == 2) {
            // ---- header exchange over NVLink: every rank -> every rank -------------------------
            const unsigned long long seq = A.xchg.epoch + (unsigned long long)iters + 1ull;
            if (cta == 0 && tid < A.world) {
                double *h = A.xchg.peer[tid] + xchg_hdr_off(A.slot_out, A.rank);
                h[0] = c.q;
                reinterpret_cast<long long *>(h)[1] = c.key;
                reinterpret_cast<long long *>(h)[2] = c.row;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long *>(
                    A.xchg.peer[tid] + xchg_flag1_off(A.slot_out, A.rank, A.ld)) = seq;
            }
            if (tid == 0) s_ok = 1;
            __syncthreads();
            if (tid < A.world) {
                const unsigned long long *f = reinterpret_cast<const unsigned long long *>(
                    A.xchg.peer[A.rank] + xchg_flag1_off(A.slot_out, tid, A.ld));
                if (!wait_flag(f, seq, A.timeout_ns)) s_ok = 0;
            }
            __syncthreads();
            if (!s_ok) {
                fin = ST_PEER_TIMEOUT;
            } else {
                if (tid == 0) {
                    __threadfence_system();
                    int ww = 0;
                    s_p = resolve_winner(A.xchg.peer[A.rank] + xchg_hdr_off(A.slot_out, 0), kCandHdr,
                                         A.world, &ww).row;
                    s_w = ww;
                }
                __syncthreads();
                win_rank = s_w;
                i_win = (win_rank == A.rank) && s_p >= 0;
                // from here on c.row is the GLOBAL leaving row (or -1: unbounded)
                if (!i_win) { c.row = s_p; }
            }
        }

        // ---- stage 3: candidate row / its pivot element -----------------------------------------
        if (fin == ST_RUNNING && i_win && c.row >= 0) {
            const int i = c.row - A.row0;
            const double s = __ldcg(col_out + i);              // written by another CTA in stage 2
            const double t = pending ? col[i] : 0.0;
            const bool is_p = (i == p_local);
            const double *rowp = A.src + (int64_t)i * A.ld;
            const int ld = (int)A.ld;
            const int chunk3 = (ld + G - 1) / G;
            const int e3 = min(ld, (cta + 1) * chunk3);
            for (int base = cta * chunk3 + tid; base < e3; base += kLookBatch * kLookThreads) {
                double vr[kLookBatch], vp[kLookBatch];
#pragma unroll
                for (int u = 0; u < kLookBatch; ++u) {
                    const int cc = base + u * kLookThreads;
                    const bool in = cc < e3 && cc < A.C;
                    vr[u] = (in && !(pending && is_p)) ? rowp[cc] : 0.0;
                    vp[u] = (in && pending) ? prow[cc] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kLookBatch; ++u) {
                    const int cc = base + u * kLookThreads;
                    if (cc >= e3) break;
                    double v = 0.0;
                    if (cc < A.C) {
                        v = vr[u];
                        if (pending) v = is_p ? vp[u] : __dsub_rn(v, __dmul_rn(t, vp[u]));
                        v = __ddiv_rn(v, s);
                    }
                    if (A.mode == 2) {
                        const int64_t off = xchg_row_off(A.slot_out, A.ld) + cc;
                        for (int g = 0; g < A.world; ++g) A.xchg.peer[g][off] = v;
                    } else {
                        cand_out[kCandHdr + cc] = v;
                    }
                }
            }
            if (A.mode == 2) __threadfence_system();
        }
    }

    // ---- leave: the last CTA publishes the verdict and re-arms the barriers --------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (fin == ST_PEER_TIMEOUT) atomicOr(&A.sync->fail, 1u);
        s_last = atomicAdd(&A.sync->bar[2], 1u) == (unsigned)(G - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) {
        __threadfence();
        if (A.sync->fail & 2u) fin = ST_BARRIER_TIMEOUT;
        else if (A.sync->fail) fin = ST_PEER_TIMEOUT;
        A.sync->bar[0] = 0; A.sync->bar[1] = 0; A.sync->bar[2] = 0; A.sync->fail = 0;
        s_ok = 1;
    }
    __syncthreads();
    if (A.mode == 2 && fin == ST_RUNNING && c.row >= 0) {
        // the winner's row is complete on every rank (all its CTAs fenced before arriving):
        // raise flag2 everywhere, then wait for ours so that the kernel ends with the row present
        const unsigned long long seq = A.xchg.epoch + (unsigned long long)iters + 1ull;
        if (win_rank == A.rank && tid < A.world) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(
                A.xchg.peer[tid] + xchg_flag2_off(A.slot_out, A.ld)) = seq;
        }
        if (tid == 0) {
            const unsigned long long *f = reinterpret_cast<const unsigned long long *>(
                A.xchg.peer[A.rank] + xchg_flag2_off(A.slot_out, A.ld));
            if (!wait_flag(f, seq, A.timeout_ns)) s_ok = 0;
            __threadfence_system();
        }
        __syncthreads();
        if (!s_ok) fin = ST_PEER_TIMEOUT;
    }
    if (tid != 0) return;
    IterState o;
    o.status = fin; o.j = j; o.p = -1; o.w = win_rank; o.iters = iters; o.pad = 0;
    if (fin == ST_RUNNING) {
        if (A.mode != 2) {
            CandHdr *hdr = reinterpret_cast<CandHdr *>(cand_out);
            hdr->q = c.q; hdr->key = c.key; hdr->row = c.row; hdr->pad = 0;
        }
        o.p = c.row;
        if (A.mode != 1 && c.row < 0) o.status = ST_UNBOUNDED;
    }
    if (o.status != ST_RUNNING) {
        A.report->iters = iters;
        A.report->status = o.status;
    }
    *st_out = o;
    A.sync->look_ns += global_timer_ns() - *reinterpret_cast<volatile unsigned long long *>(&A.sync->look_t0);
    A.sync->look_count += 1;
}

// ------------------------------------------------------------------------------------------
// update role: n-pivot-row part 2 (src/simplex.lisp:349-358) for one tile of TR rows x 256
// double2 columns, streaming src -> dst.  16*R*C algorithmic bytes per iteration in total.
// ------------------------------------------------------------------------------------------
struct UpdateArgs {
    const double *src;
    double *dst;
    int64_t ld;
    int m_local, R_local, row0, world, mode, slot;
    const IterState *ring;
    const double *colring;
    int64_t col_stride;
    const double *candring;   // mode 0
    const double *gathring;   // mode 1
    int64_t cand_stride;
    const double *xrow;       // mode 2: this rank's xchg row ring base (slot 0)
};

// Which row is the pivot row and where its scaled copy lives; false = nothing to apply.
__device__ __forceinline__ bool update_decision(const UpdateArgs &U, int *p_local, const double **prow)
{
    const IterState *st = U.ring + U.slot;
    if (st->status != ST_RUNNING) return false;
    int p = st->p, w = st->w;
    const double *row;
    if (U.mode == 1) {
        const double *base = U.gathring + (int64_t)U.slot * U.world * U.cand_stride;
        p = resolve_winner(base, U.cand_stride, U.world, &w).row;
        row = base + (int64_t)w * U.cand_stride + kCandHdr;
    } else if (U.mode == 2) {
        row = U.xrow + (int64_t)U.slot * U.ld;
    } else {
        row = U.candring + (int64_t)U.slot * U.cand_stride + kCandHdr;
    }
    if (p < 0) return false;
    const int rel = p - U.row0;
    *p_local = (rel >= 0 && rel < U.m_local) ? rel : -1;
    *prow = row;
    return true;
}

template <int TR, int UNROLL, bool STREAM>
__device__ __forceinline__ void update_tile(const double2 *src2, double2 *dst2, const int ldv,
                                            const int cv, const double2 pr, const int r_begin,
                                            const int r_end, const int p_local, const double *s_col)
{
    if (cv >= ldv) return;
    for (int r = r_begin; r < r_end; r += UNROLL) {
        double2 a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (r + u < r_end) a[u] = ld_tab<STREAM>(src2 + (int64_t)(r + u) * ldv + cv);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (r + u < r_end) {
                const double t = s_col[r + u - r_begin];
                double2 o;
                o.x = __dsub_rn(a[u].x, __dmul_rn(t, pr.x));
                o.y = __dsub_rn(a[u].y, __dmul_rn(t, pr.y));
                if (r + u == p_local) o = pr;
                st_tab<STREAM>(dst2 + (int64_t)(r + u) * ldv + cv, o);
            }
        }
    }
}

// Two-kernel packaging, update half: static 2-D grid (x: column tiles, y: row blocks).
template <int TR, int UNROLL, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads) k_update(const UpdateArgs U)
{
    __shared__ double s_col[TR];
    __shared__ int s_pl, s_go;
    __shared__ const double *s_row;
    if (threadIdx.x == 0) {
        int pl = -1;
        const double *row = nullptr;
        s_go = update_decision(U, &pl, &row) ? 1 : 0;
        s_pl = pl; s_row = row;
    }
    __syncthreads();
    if (!s_go) return;
    const int ldv = (int)(U.ld >> 1);
    const int r_begin = blockIdx.y * TR;
    const int r_end = min(U.R_local, r_begin + TR);
    const double *col = U.colring + (int64_t)U.slot * U.col_stride;
    for (int t = threadIdx.x; t < TR; t += kPivotThreads)
        s_col[t] = (r_begin + t < r_end) ? col[r_begin + t] : 0.0;
    const int cv = blockIdx.x * kPivotThreads + threadIdx.x;
    const double2 pr = (cv < ldv) ? reinterpret_cast<const double2 *>(s_row)[cv] : make_double2(0.0, 0.0);
    __syncthreads();
    update_tile<TR, UNROLL, STREAM>(reinterpret_cast<const double2 *>(U.src),
                                    reinterpret_cast<double2 *>(U.dst), ldv, cv, pr, r_begin, r_end,
                                    s_pl, s_col);
}

// Two-kernel packaging, look half.
__global__ void __launch_bounds__(kLookThreads) k_look(const LookArgs A)
{
    look_role(A, blockIdx.x, gridDim.x);
}

// One-kernel packaging: CTAs [0, look_ctas) look ahead (and exchange); CTA look_ctas + t streams
// tile t (row-block major, so consecutive CTAs cover one 64-row slab left to right).  The block
// scheduler dispatches in index order, so the look CTAs are resident before any tile.
template <int TR, int UNROLL, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads, 2)
k_iter(const LookArgs A, const UpdateArgs U, const int look_ctas)
{
    // Programmatic dependent launch: this grid may be made resident while the previous
    // iteration's grid drains; nothing it produced is touched before the wait, and the next
    // iteration's grid is allowed to queue up behind this one right away.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    if ((int)blockIdx.x < look_ctas) {
        look_role(A, blockIdx.x, look_ctas);
        return;
    }
    __shared__ double s_col[TR];
    __shared__ int s_pl, s_go;
    __shared__ const double *s_row;
    if (threadIdx.x == 0) {
        int pl = -1;
        const double *row = nullptr;
        s_go = update_decision(U, &pl, &row) ? 1 : 0;
        s_pl = pl; s_row = row;
    }
    const int ldv = (int)(U.ld >> 1);
    const int tiles_x = (ldv + kPivotThreads - 1) / kPivotThreads;
    const int tile = (int)blockIdx.x - look_ctas;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int r_begin = ty * TR;
    const int r_end = min(U.R_local, r_begin + TR);
    const double *col = U.colring + (int64_t)U.slot * U.col_stride;
    for (int t = threadIdx.x; t < TR; t += kPivotThreads)
        s_col[t] = (r_begin + t < r_end) ? col[r_begin + t] : 0.0;
    __syncthreads();
    if (!s_go) return;
    const int cv = tx * kPivotThreads + threadIdx.x;
    const double2 pr = (cv < ldv) ? reinterpret_cast<const double2 *>(s_row)[cv] : make_double2(0.0, 0.0);
    update_tile<TR, UNROLL, STREAM>(reinterpret_cast<const double2 *>(U.src),
                                    reinterpret_cast<double2 *>(U.dst), ldv, cv, pr, r_begin, r_end,
                                    s_pl, s_col);
}

// ------------------------------------------------------------------------------------------
// Small helpers for the boundary: gather the RHS column into a contiguous buffer (what
// tableau-variable reads, src/simplex.lisp:81-107) and the two-phase transition pieces.
// ------------------------------------------------------------------------------------------
__global__ void k_gather_col(const double *__restrict__ tab, int64_t ld, int R_local, int col,
                             double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R_local) out[i] = tab[(int64_t)i * ld + col];
}

// Upload staging: a host tableau with ld == C crosses PCIe as ONE contiguous copy into the spare
// ping-pong buffer and is re-pitched here to the padded device row stride (pads zeroed).
__global__ void k_repitch(const double *__restrict__ src, int64_t C, double *__restrict__ dst,
                          int64_t ld, int R_local)
{
    const int64_t n = (int64_t)R_local * ld;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = k / ld, c = k - r * ld;
        dst[k] = c < C ? src[r * C + c] : 0.0;
    }
}

__global__ void k_zero_pad(double *__restrict__ tab, int64_t ld, int R_local, int C)
{
    const int pad = (int)ld - C;
    const int64_t n = (int64_t)R_local * pad;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x)
        tab[(k / pad) * ld + C + (k % pad)] = 0.0;
}

// Phase-1 -> phase-2 coefficient copy (src/simplex.lisp:437-441): constraint rows only;
// main[r, 0..nv) = art[r, 0..nv); main[r, nv] = art[r, art_nv].
__global__ void k_copy_art_to_main(const double *__restrict__ art, int64_t ld_art, int art_nv,
                                   double *__restrict__ mtab, int64_t ld, int nv, int m)
{
    for (int r = blockIdx.y; r < m; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= nv; c += gridDim.x * blockDim.x)
            mtab[(int64_t)r * ld + c] = art[(int64_t)r * ld_art + (c < nv ? c : art_nv)];
}

// Objective re-pricing (src/simplex.lisp:444-451): for i = 0..m-1 in order, scale_i =
// obj[basis_i] (the basis columns are exact unit vectors, so the sequential read equals the
// snapshot taken here), obj[c] -= scale_i * main[i,c] when scale_i /= 0.  One thread per column
// keeps the row order, hence the rounding sequence, of the reference.
__global__ void k_reprice_scales(const double *__restrict__ obj, const int32_t *__restrict__ basis,
                                 int m, int nv, double *__restrict__ scales)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // a redundant row's zero-level artificial (B200LP_FEAS_SCALED) has no column here: nothing to price
    if (i < m) scales[i] = basis[i] < nv ? obj[basis[i]] : 0.0;
}

__global__ void k_reprice(double *__restrict__ mtab, int64_t ld, int m, int nv,
                          const double *__restrict__ scales)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nv) return;
    double acc = mtab[(int64_t)m * ld + c];
    for (int i = 0; i < m; ++i) {
        const double s = scales[i];
        if (s != 0.0) acc = __dsub_rn(acc, __dmul_rn(s, mtab[(int64_t)i * ld + c]));
    }
    mtab[(int64_t)m * ld + c] = acc;
}

// First non-basic column j < nv with an exactly non-zero entry in `row` (src/simplex.lisp:426-431);
// is_basic[] marks columns currently in the basis.  Single CTA; result -1 when none.
__global__ void k_first_nonzero_nonbasic(const double *__restrict__ row, int nv,
                                         const unsigned char *__restrict__ is_basic, int *out)
{
    __shared__ int best;
    if (threadIdx.x == 0) best = 0x7fffffff;
    __syncthreads();
    int mine = 0x7fffffff;
    for (int c = threadIdx.x; c < nv; c += blockDim.x)
        if (row[c] != 0.0 && !is_basic[c]) { mine = c; break; }
    atomicMin(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) *out = (best == 0x7fffffff) ? -1 : best;
}

// B200LP_FEAS_SCALED's replacement rule: the non-basic column j < nv with the largest |entry|
// above `thr` (first index on ties); result -1 when none.  Single CTA.
__global__ void k_largest_nonbasic(const double *__restrict__ row, int nv,
                                   const unsigned char *__restrict__ is_basic, double thr, int *out)
{
    __shared__ Cand red[1024 / 32];
    Cand best;
    best.q = 0.0; best.key = 0; best.row = -1;
    for (int c = threadIdx.x; c < nv; c += blockDim.x) {
        const double a = fabs(row[c]);
        if (a > thr && !is_basic[c]) {
            Cand d;
            d.q = -a; d.key = c; d.row = c;               // (-|a|, index) minimum == largest, first index
            best = cand_min(best, d);
        }
    }
    best = cand_block_min<1024>(best, red);
    if (threadIdx.x == 0) *out = best.row;
}

} // namespace b200lp
