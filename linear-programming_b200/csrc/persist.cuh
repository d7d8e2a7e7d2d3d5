// persist.cuh -- the whole simplex loop (src/simplex.lisp:455-460) as ONE cooperative kernel.
//
// k_iter (kernels.cuh) launches one grid per pivot; on a 1 025-row shard (config 3 on 8 GPUs) or
// an L2-resident tableau (config 2) the launch gap, the drain of the last wave of tiles and the
// grid barriers of the decision chain are a visible share of a 15-60 us iteration.  k_persist
// keeps one grid resident for the whole b200lp_iterate call (cudaLaunchCooperativeKernel: the
// co-residency the spin waits need is guaranteed by the launch, not by dispatch order):
//
//   look CTAs  [0, G)   decide pivot k = (entering column, leaving row, scaled pivot row) while
//                       the tile CTAs are still applying pivot k-1 -- the same lookahead algebra
//                       as k_iter (every cell of S_{k-1} the decision needs is recomputed from
//                       S_{k-2} and pivot k-1 with the update's own rounded product and rounded
//                       difference).  They keep COMPACT running copies of the objective row and
//                       of the RHS column (bit-identical to the tableau's, L2 resident), so only
//                       the pivot column gather and the pivot row read touch the streamed tableau.
//                       Two phases per pivot:  A  ratio test on column j_k (find-pivoting-row,
//                       :382-389) + pivot-column snapshot;  B  pivot row / pivot element
//                       (n-pivot-row part 1, :344-348) fused with the objective-row update and
//                       the argmin that picks j_{k+1} (find-entering-column, :362-379).
//   tile CTAs  [G, P)   n-pivot-row part 2 (:349-358): each owns a FIXED, balanced (to one row
//                       of 512 columns) share of the tableau for the whole call and streams it
//                       S_{k-1} -> S_k between the two ping-pong buffers as soon as decision k
//                       is published.  A thread only ever re-reads cells it wrote itself, so the
//                       tile role needs no grid barrier at all; CTAs drift up to an iteration
//                       apart and there is no tail.
//
// Flow control (PSync, global memory): `decided` = highest published decision; done[k&3] counts
// the tile CTAs that finished update(k).  Decision k needs update(k-2) complete (it reads
// S_{k-2}); update(k) needs decision k.  Every spin has a %globaltimer timeout that aborts the
// whole grid (-> B200LP_ERR_INTERNAL / B200LP_ERR_PEER_TIMEOUT) instead of hanging.
//
// Sharded (exchange mode 2, peer-mapped buffers over NVLink): after phase A every rank scales ITS
// OWN candidate row speculatively and pushes header + row to every rank with 16-byte stores, then
// raises one flag per destination; each rank waits for all flags ONCE, picks the same
// lexicographic (ratio, key) minimum and runs the objective-row half of phase B on the winner's
// row.  One flag wait per pivot (k_iter needs two: headers, then the winner's row).
//
// Arithmetic contract as everywhere: __ddiv_rn, __dmul_rn then __dsub_rn, strict compares,
// lowest index wins.  Bit-identical to k_iter, to the step-by-step kernels and to the oracle.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace b200lp {

constexpr int ST_SPIN_TIMEOUT = -6;      // == B200LP_ERR_INTERNAL
constexpr int kPLookMax = 16;            // look CTAs of k_persist
constexpr int kPUnits = 4;               // double2 column units a look thread loads per batch

// Loop state of the look role.  k_persist keeps it in registers for the whole call; k_iter2 runs
// one decision per launch and parks it here in between.
struct alignas(16) LookCarry {
    long long fin_at;          // decision index at which the solve ended (0 = still running)
    long long iters;           // pivots completed
    unsigned long long bar_n[2];
    int j, j_prev, p_prev, w_prev;
};

// Every hot word has its own 128-byte line: `decided` is polled by every tile CTA, done[] takes
// their atomics, bar[] the look CTAs' -- sharing a line makes each of them wait for the others.
struct alignas(128) PSync {
    alignas(128) unsigned long long decided;      // decisions 1..decided are in the ring
    alignas(128) unsigned long long done[kRing];  // done[k & 3]: tile CTAs that finished update(k), summed over uses
    alignas(128) unsigned long long bar[2];       // look-grid barriers (monotonic arrival counts)
    alignas(128) unsigned long long rebal;        // tile CTAs that reached a re-balancing point (monotonic)
    alignas(128) int abort;                       // != 0: some spin timed out; everyone leaves
    // telemetry, summed over decisions by look CTA 0: SM cycles per phase, plus (%globaltimer,
    // clock64) pairs at both ends of the call to convert them
    alignas(128) unsigned long long look_count;
    unsigned long long ns_look, ns_wait_done, ns_a, ns_b1, ns_xwait, ns_b2;
    unsigned long long gt0, clk0, gt1, clk1;
    unsigned int smid0, pad_smid;                 // the SM both pairs were read on (clock64 is per SM)
    unsigned long long dbg[8];                    // finer stamps of the lead thread (cycles, summed)
    alignas(128) Cand part[2][kPLookMax];
    // k_iter2 (one launch per pivot): what the look role carries from one launch to the next
    alignas(128) LookCarry carry;
};

// Exchange layout of the persistent loop (inside the same peer-mapped buffer k_iter uses):
//   cand[kRing][kMaxWorld] x (kCandHdr + ld) doubles   header + scaled candidate row of every rank
//   flag[kRing][kMaxWorld] u64                         "rank r's candidate for this slot has landed"
__host__ __device__ __forceinline__ int64_t px_cand_off(int slot, int r, int64_t ld) { return (int64_t)(slot * kMaxWorld + r) * (kCandHdr + ld); }
__host__ __device__ __forceinline__ int64_t px_flag_off(int slot, int r, int64_t ld) { return (int64_t)kRing * kMaxWorld * (kCandHdr + ld) + slot * kMaxWorld + r; }
__host__ __device__ __forceinline__ int64_t px_words(int64_t ld) { return (int64_t)kRing * kMaxWorld * (kCandHdr + ld) + kRing * kMaxWorld + 8; }

struct PersistArgs {
    double *tab[2];           // tab[i & 1] holds S_i, the tableau after i pivots of this call
    int64_t ld;
    int C, m_local, R_local, row0, world, rank;
    int is_max, rule, mode;   // mode 0: one shard; 2: peer-mapped exchange
    double thr_enter, thr_pivot;
    long long iters0;         // pivots completed before this call
    long long max_iters;      // absolute cap on the pivot count; 0 = unlimited
    IterState *ring;          // kRing decisions
    double *colring;          // kRing x col_stride pivot-column snapshots
    int64_t col_stride;
    double *prowring;         // mode 0: kRing x prow_stride scaled pivot rows
    int64_t prow_stride;
    double *objc;             // ld doubles: running copy of the objective row
    double *rhsc;             // R_local doubles: running copy of the RHS column
    Xchg xchg;
    int32_t *basis;
    Report *report;
    int2 *trace;
    int trace_cap;
    PSync *sync;
    unsigned long long timeout_ns;
    double *rates;                   // [grid]: measured row units per busy cycle of every CTA (kept across calls)
    unsigned long long *tile_prof;   // optional [2 * grid]: busy ns and SM id per tile CTA (B200LP_TILE_PROFILE)
    int look_ctas;
    int slot_base;            // ring slot of decision k is (slot_base + k) & 3: the ring keeps turning
                              // across calls, so a rank that is already in the next call never
                              // overwrites a slot a slower rank's last update still reads
};

// ---- spins -------------------------------------------------------------------------------------
__device__ __forceinline__ bool spin_ge(const unsigned long long *p, unsigned long long want,
                                        PSync *S, unsigned long long timeout_ns, int code)
{
    // deadline on the SM cycle counter (reading %globaltimer costs about a microsecond, which a
    // latency chain cannot afford): 2 cycles per ns, i.e. between 1x and 2x the nominal timeout
    const volatile unsigned long long *v = p;
    if (*v >= want) return true;
    const long long t0 = clock64();
    const long long budget = (long long)(timeout_ns << 1);
    for (;;) {
        for (int k = 0; k < 32; ++k)
            if (*v >= want) return true;
        if (*reinterpret_cast<volatile int *>(&S->abort)) return false;
        if (clock64() - t0 > budget) {
            atomicCAS(&S->abort, 0, code);
            return false;
        }
    }
}

// Consumer side of every flag in this file: the producer fences (gpu or system scope) between its
// data and the flag; the consumer polls the flag with volatile loads and then reads the data ONLY
// with loads that are served by L2 (ld.cg / volatile), issued after the poll returned (control
// dependency, and bar.sync for the other threads of the CTA).  A consumer-side __threadfence()
// would add about a microsecond to every hop of the decision chain (it has to wait for the
// thread's own outstanding stores) without making anything more visible than L2 already is.
//
// Thread 0 waits, everybody learns the outcome.
__device__ __forceinline__ bool cta_wait_ge(const unsigned long long *p, unsigned long long want,
                                            PSync *S, unsigned long long timeout_ns, int code)
{
    int ok = 1;
    if (threadIdx.x == 0) ok = spin_ge(p, want, S, timeout_ns, code) ? 1 : 0;
    return __syncthreads_and(ok) != 0;
}

// Barrier of the G look CTAs (co-resident: cooperative launch).  `target` = G x (uses so far).
// `sys`: the writes being published went to peer GPUs (system scope).  One fence by thread 0 after
// the CTA barrier covers the whole CTA's writes (cumulativity) -- a membar.sys per thread, thousands
// per pivot, measurably slows the update tiles running beside the look CTAs.
__device__ __forceinline__ bool look_bar(PSync *S, int which, unsigned long long target,
                                         unsigned long long timeout_ns, const bool sys = false)
{
    __syncthreads();
    int ok = 1;
    if (threadIdx.x == 0) {
        if (sys) __threadfence_system();
        else __threadfence();
        atomicAdd(&S->bar[which], 1ull);
        ok = spin_ge(&S->bar[which], target, S, timeout_ns, ST_SPIN_TIMEOUT) ? 1 : 0;
    }
    return __syncthreads_and(ok) != 0;
}

__device__ __forceinline__ Cand preduce(const Cand *part, int G, Cand *s_out)
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

// The look grid as ONE thread-block cluster (k_persist, one shard): the partial argmins travel
// through distributed shared memory and the CTAs meet at the hardware cluster barrier
// (barrier.cluster, release/acquire at cluster scope) instead of a fence + atomic + spin on a
// global counter.  Every CTA reads all G partials from its peers' shared memory and reduces them
// in the same order, so all agree.  `mine` is this CTA's slot; A and B phases use different slots,
// so a slot is rewritten only after a later cluster barrier that everybody's read precedes.
__device__ __forceinline__ Cand cluster_reduce(Cand *mine, const Cand v, const int G, Cand *s_out)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) *mine = v;
    cluster.sync();
    if (threadIdx.x < 32) {
        Cand d;
        d.q = 0.0; d.key = 0; d.row = -1;
        if ((int)threadIdx.x < G) {
            const Cand *r = cluster.map_shared_rank(mine, threadIdx.x);
            d = *r;
        }
        d = cand_warp_min(d);
        if (threadIdx.x == 0) *s_out = d;
    }
    __syncthreads();
    return *s_out;
}

__device__ __forceinline__ const double *p_prow(const PersistArgs &P, int slot, int w)
{
    if (P.mode == 2) return P.xchg.peer[P.rank] + px_cand_off(slot, w, P.ld) + kCandHdr;
    return P.prowring + (int64_t)slot * P.prow_stride;
}

__device__ __forceinline__ double2 ldcg2(const double *p)
{
    return __ldcg(reinterpret_cast<const double2 *>(p));
}
__device__ __forceinline__ void stcg2(double *p, const double2 v)
{
    __stcg(reinterpret_cast<double2 *>(p), v);
}

// argmin bookkeeping of find-entering-column over one objective-row value
__device__ __forceinline__ void enter_scan(Cand &best, const double o, const int cc, const int nv,
                                           const int is_max, const int rule, const double thr)
{
    if (cc >= nv) return;
    const double key = is_max ? o : -o;
    if (rule == 0) {
        // thread-local order is ascending in cc, so strict < keeps the first extremum
        if (best.row < 0 || key < best.q) { best.q = key; best.key = cc; best.row = cc; }
    } else if (best.row < 0 && key < 0.0 - thr) {
        best.q = 0.0; best.key = cc; best.row = cc;              // Bland: lowest passing index
    }
}

// ---- look role ---------------------------------------------------------------------------------
// ONE_STEP = false: the whole loop (k_persist).  ONE_STEP = true: decision `k_one` only (k_iter2:
// launch L decides pivot L + 1; the previous launch is complete, so there is nothing to wait for);
// the loop state comes from / goes back to PSync::carry.
template <bool ONE_STEP, bool CLUSTER = false>
__device__ __forceinline__ void persist_look(const PersistArgs &P, const int cta, const int G,
                                             const long long k_one)
{
    __shared__ Cand s_cpart[2];                            // CLUSTER: this CTA's partials (phase A, phase B)
    __shared__ Cand red[kLookThreads / 32];
    __shared__ Cand s_part;
    __shared__ int s_p, s_w;
    const int tid = threadIdx.x;
    PSync *S = P.sync;
    const int nv = P.C - 1, rhs = P.C - 1;
    const int ld = (int)P.ld, ldv = ld >> 1;
    const bool lead = (cta == 0 && tid == 0);
    const unsigned long long T = gridDim.x - G;
    unsigned long long bar_n[2] = {0ull, 0ull};
    const int chunkV = (ldv + G - 1) / G;
    const int v_begin = cta * chunkV, v_end = min(ldv, v_begin + chunkV);
    const int chunkR = (P.R_local + G - 1) / G;
    const int r_begin = cta * chunkR, r_end = min(P.R_local, r_begin + chunkR);
    long long iters = P.iters0;
    int j = -1;
    int j_prev = -1, p_prev = -1, w_prev = 0;
    Cand best;
    best.q = 0.0; best.key = 0; best.row = -1;
    if (ONE_STEP && k_one > 1) {
        const LookCarry *c = &S->carry;
        iters = __ldcg(&c->iters);
        bar_n[0] = __ldcg(&c->bar_n[0]); bar_n[1] = __ldcg(&c->bar_n[1]);
        j = __ldcg(&c->j); j_prev = __ldcg(&c->j_prev);
        p_prev = __ldcg(&c->p_prev); w_prev = __ldcg(&c->w_prev);
    } else {
    if (lead) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        S->smid0 = smid;
        S->gt0 = global_timer_ns();
        S->clk0 = (unsigned long long)clock64();
    }

    // ---- prologue: compact copies of S_0's objective row / RHS column; entering column of pivot 1
    {
        const double *S0 = P.tab[0];
        const double *obj = S0 + (int64_t)P.m_local * P.ld;
        for (int v = v_begin + tid; v < v_end; v += kLookThreads) {
            const double2 o = ldcg2(obj + 2 * v);
            stcg2(P.objc + 2 * v, o);
            enter_scan(best, o.x, 2 * v, nv, P.is_max, P.rule, P.thr_enter);
            enter_scan(best, o.y, 2 * v + 1, nv, P.is_max, P.rule, P.thr_enter);
        }
        for (int i = r_begin + tid; i < r_end; i += kLookThreads)
            __stcg(P.rhsc + i, __ldcg(S0 + (int64_t)i * P.ld + rhs));
    }
    best = cand_block_min<kLookThreads>(best, red);
    if (G > 1) {
        if (CLUSTER) {
            best = cluster_reduce(&s_cpart[1], best, G, &s_part);
        } else {
            if (tid == 0) S->part[1][cta] = best;
            bar_n[1] += G;
            if (!look_bar(S, 1, bar_n[1], P.timeout_ns)) return;
            best = preduce(S->part[1], G, &s_part);
        }
    }
    {
        const bool accept = (best.row >= 0) && (P.rule != 0 || best.q < 0.0 - P.thr_enter);
        int fin = ST_RUNNING;
        if (!accept) fin = ST_OPTIMAL;
        else if (P.max_iters > 0 && iters >= P.max_iters) fin = ST_ITERATION_LIMIT;
        j = accept ? best.row : -1;
        if (fin != ST_RUNNING) {
            if (lead) {
                IterState o;
                o.status = fin; o.j = j; o.p = -1; o.w = 0; o.iters = iters; o.pad = 0;
                P.ring[(1 + P.slot_base) & (kRing - 1)] = o;
                P.report->iters = iters;
                P.report->status = fin;
                S->carry.fin_at = 1;
                __threadfence();
                *reinterpret_cast<volatile unsigned long long *>(&S->decided) = 1ull;
            }
            return;
        }
    }
    }   // prologue

    for (long long k = ONE_STEP ? k_one : 1;; ++k) {
        const int slot = (int)((k + P.slot_base) & (kRing - 1));
        const int pslot = (int)((k - 1 + P.slot_base) & (kRing - 1));
        const bool pending = k >= 2;
        const double *src = P.tab[pending ? (int)(k & 1) : 0];       // S_{k-2} (S_0 for k = 1)
        const long long t0 = lead ? clock64() : 0ll;
        if (!ONE_STEP && k >= 3) {
            // S_{k-2} must be complete: every tile CTA has finished update(k-2)
            const unsigned long long want = T * (unsigned long long)((k - 3) / kRing + 1);
            if (!cta_wait_ge(&S->done[(k - 2) & (kRing - 1)], want, S, P.timeout_ns, ST_SPIN_TIMEOUT))
                return;
        }
        const long long t1 = lead ? clock64() : 0ll;
        const double *colp = P.colring + (int64_t)pslot * P.col_stride;
        const double *prowp = p_prow(P, pslot, w_prev);
        double *col_out = P.colring + (int64_t)slot * P.col_stride;
        int pp_local = -1;
        if (pending) {
            const int rel = p_prev - P.row0;
            if (rel >= 0 && rel < P.m_local) pp_local = rel;
        }

        // ---- phase A: pivot-column snapshot + ratio test over this shard's rows ---------------
        Cand c;
        c.q = 0.0; c.key = 0; c.row = -1;
        {
            const double pj = pending ? __ldcg(prowp + j) : 0.0;
            const double pb = pending ? __ldcg(prowp + rhs) : 0.0;
            for (int base = r_begin + tid; base < r_end; base += kLookRows * kLookThreads) {
                double va[kLookRows], vb[kLookRows], vt[kLookRows];
#pragma unroll
                for (int u = 0; u < kLookRows; ++u) {
                    const int i = base + u * kLookThreads;
                    const bool in = i < r_end;
                    va[u] = in ? __ldcg(src + (int64_t)i * P.ld + j) : 0.0;
                    vb[u] = in ? __ldcg(P.rhsc + i) : 0.0;
                    vt[u] = (pending && in) ? __ldcg(colp + i) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kLookRows; ++u) {
                    const int i = base + u * kLookThreads;
                    if (i >= r_end) break;
                    double a = va[u], b = vb[u];
                    if (pending) {
                        const double t = vt[u];
                        a = (i == pp_local) ? pj : __dsub_rn(a, __dmul_rn(t, pj));
                        b = (i == pp_local) ? pb : __dsub_rn(b, __dmul_rn(t, pb));
                        __stcg(P.rhsc + i, b);                     // RHS column of S_{k-1}
                    }
                    __stcg(col_out + i, a);
                    if (i < P.m_local && 0.0 + P.thr_pivot < a) {
                        Cand d;
                        d.q = __ddiv_rn(b, a);
                        d.key = P.rule ? ((i == pp_local) ? j_prev : __ldcg(P.basis + i)) : P.row0 + i;
                        d.row = P.row0 + i;
                        c = cand_min(c, d);
                    }
                }
            }
        }
        const long long ta1 = lead ? clock64() : 0ll;
        __syncthreads();                                           // red[] reuse
        c = cand_block_min<kLookThreads>(c, red);
        if (G > 1) {
            if (CLUSTER) {
                c = cluster_reduce(&s_cpart[0], c, G, &s_part);
            } else {
                if (tid == 0) S->part[0][cta] = c;
                bar_n[0] += G;
                if (!look_bar(S, 0, bar_n[0], P.timeout_ns)) return;
                c = preduce(S->part[0], G, &s_part);
            }
        }
        const long long t2 = lead ? clock64() : 0ll;
        long long t3 = t2, t4 = t2;

        // ---- phase B: candidate row / pivot element, objective-row update, next entering column
        int p = c.row, w = 0;
        best.q = 0.0; best.key = 0; best.row = -1;
        const double tobj = __ldcg(col_out + P.m_local);           // objective-row entry of column j
        if (P.mode != 2) {
            if (p >= 0) {
                const int i = p - P.row0;
                const double s = __ldcg(col_out + i);
                const double t = pending ? __ldcg(colp + i) : 0.0;
                const bool is_pp = pending && (i == pp_local);
                const double *rowp = src + (int64_t)i * P.ld;
                double *prow_out = P.prowring + (int64_t)slot * P.prow_stride;
                for (int base = v_begin + tid; base < v_end; base += kPUnits * kLookThreads) {
                    double2 vr[kPUnits], vp[kPUnits], vo[kPUnits];
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        const bool in = v < v_end;
                        vr[u] = (in && !is_pp) ? ldcg2(rowp + 2 * v) : make_double2(0.0, 0.0);
                        vp[u] = (in && pending) ? ldcg2(prowp + 2 * v) : make_double2(0.0, 0.0);
                        vo[u] = in ? ldcg2(P.objc + 2 * v) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        if (v >= v_end) break;
                        double x0 = vr[u].x, x1 = vr[u].y;
                        if (pending) {
                            x0 = is_pp ? vp[u].x : __dsub_rn(x0, __dmul_rn(t, vp[u].x));
                            x1 = is_pp ? vp[u].y : __dsub_rn(x1, __dmul_rn(t, vp[u].y));
                        }
                        x0 = (2 * v < P.C) ? __ddiv_rn(x0, s) : 0.0;
                        x1 = (2 * v + 1 < P.C) ? __ddiv_rn(x1, s) : 0.0;
                        double2 o;
                        o.x = __dsub_rn(vo[u].x, __dmul_rn(tobj, x0));
                        o.y = __dsub_rn(vo[u].y, __dmul_rn(tobj, x1));
                        stcg2(prow_out + 2 * v, make_double2(x0, x1));
                        stcg2(P.objc + 2 * v, o);
                        enter_scan(best, o.x, 2 * v, nv, P.is_max, P.rule, P.thr_enter);
                        enter_scan(best, o.y, 2 * v + 1, nv, P.is_max, P.rule, P.thr_enter);
                    }
                }
            }
        } else {
            // B1: scale OUR candidate row and push header + row to every rank (one flag each)
            const unsigned long long seq = P.xchg.epoch + (unsigned long long)k;
            if (p >= 0) {
                const int i = p - P.row0;
                const double s = __ldcg(col_out + i);
                const double t = pending ? __ldcg(colp + i) : 0.0;
                const bool is_pp = pending && (i == pp_local);
                const double *rowp = src + (int64_t)i * P.ld;
                const int64_t off = px_cand_off(slot, P.rank, P.ld) + kCandHdr;
                for (int base = v_begin + tid; base < v_end; base += kPUnits * kLookThreads) {
                    double2 vr[kPUnits], vp[kPUnits];
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        const bool in = v < v_end;
                        vr[u] = (in && !is_pp) ? ldcg2(rowp + 2 * v) : make_double2(0.0, 0.0);
                        vp[u] = (in && pending) ? ldcg2(prowp + 2 * v) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        if (v >= v_end) break;
                        double x0 = vr[u].x, x1 = vr[u].y;
                        if (pending) {
                            x0 = is_pp ? vp[u].x : __dsub_rn(x0, __dmul_rn(t, vp[u].x));
                            x1 = is_pp ? vp[u].y : __dsub_rn(x1, __dmul_rn(t, vp[u].y));
                        }
                        x0 = (2 * v < P.C) ? __ddiv_rn(x0, s) : 0.0;
                        x1 = (2 * v + 1 < P.C) ? __ddiv_rn(x1, s) : 0.0;
                        const double2 x = make_double2(x0, x1);
                        for (int g = 0; g < P.world; ++g)
                            *reinterpret_cast<double2 *>(P.xchg.peer[g] + off + 2 * v) = x;
                    }
                }
            }
            // the header goes out with the row; the barrier's system fence (one per CTA, after the
            // CTA barrier) publishes both, and only the flag is stored behind it -- one membar.sys
            // on the chain, not two
            if (cta == 0 && tid < P.world) {
                double *h = P.xchg.peer[tid] + px_cand_off(slot, P.rank, P.ld);
                h[0] = c.q;
                reinterpret_cast<long long *>(h)[1] = c.key;
                reinterpret_cast<long long *>(h)[2] = c.row;
            }
            if (G > 1) {                                           // every CTA's slice is on its way
                bar_n[1] += G;
                if (!look_bar(S, 1, bar_n[1], P.timeout_ns, /*sys=*/true)) return;
            } else {
                __syncthreads();
                if (tid == 0) __threadfence_system();
                __syncthreads();
            }
            if (cta == 0 && tid < P.world)
                *reinterpret_cast<volatile unsigned long long *>(
                    P.xchg.peer[tid] + px_flag_off(slot, P.rank, P.ld)) = seq;
            t3 = lead ? clock64() : 0ll;
            // B2: ONE wait for every rank's candidate, then the same winner everywhere
            {
                int ok = 1;
                if (tid < P.world) {
                    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(
                        P.xchg.peer[P.rank] + px_flag_off(slot, tid, P.ld));
                    ok = spin_ge(f, seq, S, P.timeout_ns, ST_PEER_TIMEOUT) ? 1 : 0;
                }
                if (!__syncthreads_and(ok)) return;
            }
            if (tid == 0) {
                int ww = 0;
                s_p = resolve_winner(P.xchg.peer[P.rank] + px_cand_off(slot, 0, P.ld),
                                     kCandHdr + P.ld, P.world, &ww).row;
                s_w = ww;
            }
            __syncthreads();
            p = s_p; w = s_w;
            t4 = lead ? clock64() : 0ll;
            if (p >= 0) {
                const double *prow_k = p_prow(P, slot, w);
                for (int base = v_begin + tid; base < v_end; base += kPUnits * kLookThreads) {
                    double2 vp[kPUnits], vo[kPUnits];
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        const bool in = v < v_end;
                        vp[u] = in ? ldcg2(prow_k + 2 * v) : make_double2(0.0, 0.0);
                        vo[u] = in ? ldcg2(P.objc + 2 * v) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < kPUnits; ++u) {
                        const int v = base + u * kLookThreads;
                        if (v >= v_end) break;
                        double2 o;
                        o.x = __dsub_rn(vo[u].x, __dmul_rn(tobj, vp[u].x));
                        o.y = __dsub_rn(vo[u].y, __dmul_rn(tobj, vp[u].y));
                        stcg2(P.objc + 2 * v, o);
                        enter_scan(best, o.x, 2 * v, nv, P.is_max, P.rule, P.thr_enter);
                        enter_scan(best, o.y, 2 * v + 1, nv, P.is_max, P.rule, P.thr_enter);
                    }
                }
            }
        }
        if (p < 0) {                                               // no eligible row anywhere: unbounded
            if (lead) {
                IterState o;
                o.status = ST_UNBOUNDED; o.j = j; o.p = -1; o.w = 0; o.iters = iters; o.pad = 0;
                P.ring[slot] = o;
                P.report->iters = iters;
                P.report->status = ST_UNBOUNDED;
                S->carry.fin_at = k;
                __threadfence();
                *reinterpret_cast<volatile unsigned long long *>(&S->decided) = (unsigned long long)k;
            }
            return;
        }
        const long long tb1 = lead ? clock64() : 0ll;
        __syncthreads();                                           // red[] reuse
        best = cand_block_min<kLookThreads>(best, red);
        if (G > 1) {
            if (CLUSTER) {
                best = cluster_reduce(&s_cpart[1], best, G, &s_part);
            } else {
                if (tid == 0) S->part[1][cta] = best;
                bar_n[1] += G;
                if (!look_bar(S, 1, bar_n[1], P.timeout_ns)) return;
                best = preduce(S->part[1], G, &s_part);
            }
        }
        const long long tb2 = lead ? clock64() : 0ll;
        // the scaled pivot row is complete (every look CTA passed the barrier): publish pivot k
        const long long iters_after = iters + 1;
        const bool accept = (best.row >= 0) && (P.rule != 0 || best.q < 0.0 - P.thr_enter);
        int fin = ST_RUNNING;
        if (!accept) fin = ST_OPTIMAL;
        else if (P.max_iters > 0 && iters_after >= P.max_iters) fin = ST_ITERATION_LIMIT;
        const int jn = accept ? best.row : -1;
        if (lead) {
            const int rel = p - P.row0;
            if (rel >= 0 && rel < P.m_local) P.basis[rel] = j;      // (setf (aref basis p) j) :358
            if (P.trace && iters < P.trace_cap) P.trace[iters] = make_int2(j, p);
            IterState o;
            o.status = ST_RUNNING; o.j = j; o.p = p; o.w = w; o.iters = iters; o.pad = 0;
            P.ring[slot] = o;
            if (fin != ST_RUNNING) {
                IterState e;
                e.status = fin; e.j = jn; e.p = -1; e.w = 0; e.iters = iters_after; e.pad = 0;
                P.ring[(k + 1 + P.slot_base) & (kRing - 1)] = e;
                P.report->iters = iters_after;
                P.report->status = fin;
                S->carry.fin_at = k + 1;
            }
            if (ONE_STEP) {
                LookCarry *c = &S->carry;
                c->iters = iters_after;
                c->bar_n[0] = bar_n[0]; c->bar_n[1] = bar_n[1];
                c->j = jn; c->j_prev = j; c->p_prev = p; c->w_prev = w;
            }
            __threadfence();
            *reinterpret_cast<volatile unsigned long long *>(&S->decided) =
                (unsigned long long)(fin != ST_RUNNING ? k + 1 : k);
            const long long t5 = clock64();
            S->look_count += 1;
            S->ns_wait_done += (unsigned long long)(t1 - t0);
            S->ns_a += (unsigned long long)(t2 - t1);
            S->ns_b1 += (unsigned long long)(t3 - t2);
            S->ns_xwait += (unsigned long long)(t4 - t3);
            S->ns_b2 += (unsigned long long)(t5 - t4);
            S->dbg[0] += (unsigned long long)(ta1 - t1);     // phase A: loads + math + stores issued
            S->dbg[1] += (unsigned long long)(t2 - ta1);     // phase A: reduce (+ look-grid barrier)
            S->dbg[2] += (unsigned long long)(tb1 - t4);     // phase B: loads + division + stores issued
            S->dbg[3] += (unsigned long long)(tb2 - tb1);    // phase B: reduce (+ look-grid barrier)
            S->dbg[4] += (unsigned long long)(t5 - tb2);     // publication (fence)
            // second (%globaltimer, clock64) pair for the cycles -> ns conversion: clock64 is a
            // per-SM counter, so it must come from the SM the first pair was read on (k_persist:
            // always; k_iter2: whenever a later launch's lead CTA lands there again)
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            if (smid == S->smid0 && (fin != ST_RUNNING || ONE_STEP)) {
                S->gt1 = global_timer_ns();
                S->clk1 = (unsigned long long)clock64();
            }
        }
        if (fin != ST_RUNNING || ONE_STEP) return;
        j_prev = j; p_prev = p; w_prev = w;
        j = jn;
        iters = iters_after;
    }
}

// ---- tile role -----------------------------------------------------------------------------------
// Work space: the tableau cut into tiles of kPTileRows rows x 256 double2 columns, numbered
// slab-major (all column tiles of rows 0..15 left to right, then rows 16..31, ...), counted in
// ROW UNITS (one row of one tile).  Tile CTA c owns the contiguous run [ru0, ru1) of that space
// for as long as the partition stands, so a thread only ever re-reads cells it wrote itself and
// the tile role needs no grid barrier between pivots.
//
// Equal shares are not equal times: on B200 an SM's share of the memory system depends on where
// it sits (a fixed-share run measured 377..562 us per CTA on config 3 -- the slowest CTA sets the
// pace; the per-pivot k_iter launch gets its balance from the hardware CTA scheduler instead).
// So shares are proportional to each CTA's MEASURED rate (row units per busy cycle): `rates`
// lives with the shard across calls, and inside a call the tile CTAs re-balance at pivots
// 4, 16, 64, ... and every 4096 after that -- a barrier among the tile CTAs only (ownership may
// move only when every write of the previous pivot is visible), off the look CTAs' chain.
constexpr int kPTileRows = 16;

struct PTile { int tx, ra, rb; };

struct PWork {
    int tiles_x, R_local;
    int rl;                         // rows of the last slab (kPTileRows when it is a full one)
    int RU;                         // row units in the full 16-row slabs: what the shares divide
};

__device__ __forceinline__ PWork pwork_make(const int ldv, const int R_local)
{
    PWork q;
    q.R_local = R_local;
    q.tiles_x = (ldv + kPivotThreads - 1) / kPivotThreads;
    const int tiles_y = (R_local + kPTileRows - 1) / kPTileRows;
    q.rl = R_local - (tiles_y - 1) * kPTileRows;
    // <= R * C / 512 < 2^31 for anything that fits in HBM
    q.RU = (q.rl == kPTileRows ? tiles_y : tiles_y - 1) * q.tiles_x * kPTileRows;
    return q;
}

// Next piece of this CTA's work: first its run [ru, ru1) of the full slabs, one tile (or the part
// of a tile inside the run) at a time; then its tiles of the short last slab, which are dealt
// round-robin (tile tx to CTA tx mod T) -- a run of 1-row pieces would keep one load per thread
// in flight and make its owner the slowest CTA (config 3 on 8 GPUs: 1 025 rows, last slab = the
// objective row alone).  false when nothing is left.
__device__ __forceinline__ bool ptile_next(const PWork &q, int &ru, const int ru1, int &xt,
                                           const int T, PTile &o)
{
    if (ru < ru1) {
        const int t = ru / kPTileRows, rr = ru % kPTileRows;
        const int take = min(ru1 - ru, kPTileRows - rr);
        const int ty = t / q.tiles_x;
        o.tx = t - ty * q.tiles_x;
        o.ra = ty * kPTileRows + rr;
        o.rb = o.ra + take;
        ru += take;
        return true;
    }
    if (q.rl < kPTileRows && xt < q.tiles_x) {
        o.tx = xt;
        o.ra = q.R_local - q.rl;
        o.rb = q.R_local;
        xt += T;
        return true;
    }
    return false;
}

// Share of tile CTA `tcta` under the current rate table.  All threads stage the table in shared
// memory, thread 0 sums it.  Every CTA evaluates the same expression on the same table in the
// same order, so neighbouring shares meet exactly.  Ends with a __syncthreads.
constexpr int kPMaxCtas = 1024;
__device__ __forceinline__ void ppart_shares(const double *rates, const int G, const int T,
                                             const int tcta, const int RU, double *s_rates, int *s_ru)
{
    for (int c = threadIdx.x; c < T; c += kPivotThreads) s_rates[c] = __ldcg(rates + G + c);
    __syncthreads();
    if (threadIdx.x == 0) {
        double mean = 0.0;
        for (int c = 0; c < T; ++c) mean += (s_rates[c] > 0.0) ? s_rates[c] : 0.0;
        mean = mean > 0.0 ? mean / T : 1.0;
        const double lo = 0.5 * mean, hi = 2.0 * mean;
        double tot = 0.0, before = 0.0, upto = 0.0;
        for (int c = 0; c < T; ++c) {
            double r = s_rates[c];
            r = (r > 0.0) ? fmin(fmax(r, lo), hi) : mean;
            if (c == tcta) before = tot;
            tot += r;
            if (c == tcta) upto = tot;
        }
        s_ru[0] = (tcta == 0) ? 0 : (int)((double)RU * (before / tot));
        s_ru[1] = (tcta == T - 1) ? RU : (int)((double)RU * (upto / tot));
    }
    __syncthreads();
}

// n-pivot-row part 2 for rows [ra, rb) (at most kPTileRows) of one column tile.  `cl` is this
// lane's pivot-column value (lane u holds col[ra + u]), `pr` the thread's slice of the pivot row.
template <bool STREAM>
__device__ __forceinline__ void persist_rows(const double2 *src2, double2 *dst2, const int ldv,
                                             const int cv, const bool act, const int ra, const int rb,
                                             const double cl, const double2 pr, const int p_local)
{
    double2 a[kPTileRows];
#pragma unroll
    for (int u = 0; u < kPTileRows; ++u)
        if (act && ra + u < rb) a[u] = ld_tab<STREAM>(src2 + (int64_t)(ra + u) * ldv + cv);
#pragma unroll
    for (int u = 0; u < kPTileRows; ++u) {
        const double t = __shfl_sync(0xffffffffu, cl, u);
        if (act && ra + u < rb) {
            double2 o;
            o.x = __dsub_rn(a[u].x, __dmul_rn(t, pr.x));
            o.y = __dsub_rn(a[u].y, __dmul_rn(t, pr.y));
            if (ra + u == p_local) o = pr;
            st_tab<STREAM>(dst2 + (int64_t)(ra + u) * ldv + cv, o);
        }
    }
}

__device__ __forceinline__ bool prebalance_at(const long long k)
{
    if (k < 4) return false;
    if ((k & 4095) == 0) return true;
    return k <= 1024 && (k & (k - 1)) == 0 && (__ffsll(k) & 1) == 1;   // 4, 16, 64, 256, 1024
}

template <int UNROLL, bool STREAM>
__device__ __forceinline__ void persist_tiles(const PersistArgs &P, const int tcta, const int T)
{
    __shared__ int s_status, s_pw[2], s_ru[2];
    __shared__ double s_rates[kPMaxCtas];
    PSync *S = P.sync;
    const int ldv = (int)(P.ld >> 1);
    const int lane = threadIdx.x & 31;
    const int G = (int)gridDim.x - T;
    const PWork q = pwork_make(ldv, P.R_local);
    // enough work for shares to matter
    const bool balance = P.rates != nullptr && q.RU >= 64 * T && T <= kPMaxCtas;
    if (balance) ppart_shares(P.rates, G, T, tcta, q.RU, s_rates, s_ru);
    if (threadIdx.x == 0) {
        if (!balance) {
            s_ru[0] = (int)((long long)q.RU * tcta / T);
            s_ru[1] = (int)((long long)q.RU * (tcta + 1) / T);
        }
        if (P.tile_prof) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.tile_prof[2 * blockIdx.x + 1] = smid;
        }
    }
    __syncthreads();
    int ru0 = s_ru[0], ru1 = s_ru[1];
    long long busy = 0, busy_all = 0;        // thread 0: cycles streaming since the last re-balance / in all
    long long units = 0;
    unsigned long long rebal_gen = 0;
    for (long long k = 1;; ++k) {
        const int slot = (int)((k + P.slot_base) & (kRing - 1));
        int ok = 1;
        if (threadIdx.x == 0) {
            ok = spin_ge(&S->decided, (unsigned long long)k, S, P.timeout_ns, ST_SPIN_TIMEOUT) ? 1 : 0;
            const volatile IterState *st = P.ring + slot;
            s_status = st->status;
            s_pw[0] = st->p;
            s_pw[1] = st->w;
        }
        if (!__syncthreads_and(ok)) return;
        if (s_status != ST_RUNNING) return;
        const int p = s_pw[0], w = s_pw[1];
        if (balance && prebalance_at(k)) {
            // every tile CTA has finished update(k-1) and published its rate before any of them
            // starts update(k) under the new shares
            rebal_gen += (unsigned long long)T;
            if (threadIdx.x == 0) {
                if (busy > 0 && units > 0) P.rates[blockIdx.x] = (double)units / (double)busy;
                __threadfence();
                atomicAdd(&S->rebal, 1ull);
                ok = spin_ge(&S->rebal, rebal_gen, S, P.timeout_ns, ST_SPIN_TIMEOUT) ? 1 : 0;
                busy = 0; units = 0;
            }
            if (!__syncthreads_and(ok)) return;
            ppart_shares(P.rates, G, T, tcta, q.RU, s_rates, s_ru);
            ru0 = s_ru[0]; ru1 = s_ru[1];
        }
        const int rel = p - P.row0;
        const int p_local = (rel >= 0 && rel < P.m_local) ? rel : -1;
        const double2 *src2 = reinterpret_cast<const double2 *>(P.tab[(k - 1) & 1]);
        double2 *dst2 = reinterpret_cast<double2 *>(P.tab[k & 1]);
        const double *col = P.colring + (int64_t)slot * P.col_stride;
        const double *prow = p_prow(P, slot, w);
        const long long tb = (threadIdx.x == 0) ? clock64() : 0ll;
        // the pivot-row slice and the pivot-column values of the NEXT piece are fetched while
        // the current one streams
        int ru = ru0, xt = tcta;
        PTile nx;
        nx.tx = 0; nx.ra = 0; nx.rb = 0;
        bool have = ptile_next(q, ru, ru1, xt, T, nx);
        double2 pr_n = make_double2(0.0, 0.0);
        double cl_n = 0.0;
        if (have) {
            const int cvn = nx.tx * kPivotThreads + (int)threadIdx.x;
            if (cvn < ldv) pr_n = ldcg2(prow + 2 * cvn);
            if (lane < kPTileRows && nx.ra + lane < nx.rb) cl_n = __ldcg(col + nx.ra + lane);
        }
        while (have) {
            const PTile cur = nx;
            const double2 pr = pr_n;
            const double cl = cl_n;
            have = ptile_next(q, ru, ru1, xt, T, nx);
            if (have) {
                const int cvn = nx.tx * kPivotThreads + (int)threadIdx.x;
                pr_n = (cvn < ldv) ? ldcg2(prow + 2 * cvn) : make_double2(0.0, 0.0);
                cl_n = (lane < kPTileRows && nx.ra + lane < nx.rb) ? __ldcg(col + nx.ra + lane) : 0.0;
            }
            const int cv = cur.tx * kPivotThreads + (int)threadIdx.x;
            persist_rows<STREAM>(src2, dst2, ldv, cv, cv < ldv, cur.ra, cur.rb, cl, pr, p_local);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long dt = clock64() - tb;
            busy += dt; busy_all += dt;
            units += ru1 - ru0;
            __threadfence();
            atomicAdd(&S->done[k & (kRing - 1)], 1ull);
            if (P.tile_prof) P.tile_prof[2 * blockIdx.x] = (unsigned long long)busy_all;
        }
    }
}

template <int UNROLL, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads, 2) k_persist(const __grid_constant__ PersistArgs P)
{
    const int G = P.look_ctas;
    if ((int)blockIdx.x < G) persist_look<false>(P, blockIdx.x, G, 1);
    else persist_tiles<UNROLL, STREAM>(P, (int)blockIdx.x - G, (int)gridDim.x - G);
}

// The same kernel launched with thread-block clusters of G CTAs (cooperative + cluster launch):
// cluster 0 is the look grid and synchronises through barrier.cluster / distributed shared
// memory; the tile CTAs sit in clusters too but never use them.  One shard only (mode 0): the
// hardware barrier has no timeout, so it is kept away from waits on other GPUs.
template <int UNROLL, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads, 2) k_persist_cl(const __grid_constant__ PersistArgs P)
{
    const int G = P.look_ctas;
    if ((int)blockIdx.x < G) persist_look<false, true>(P, blockIdx.x, G, 1);
    else persist_tiles<UNROLL, STREAM>(P, (int)blockIdx.x - G, (int)gridDim.x - G);
}

// ------------------------------------------------------------------------------------------------
// k_iter2: the per-pivot packaging (one launch per pivot, PDL) with the look role of this file:
// launch L applies pivot L (update tiles, one CTA per 16-row tile -- the hardware CTA scheduler
// balances the SMs) while its look CTAs take decision L + 1 in two phases on the compact
// objective-row / RHS copies.  Sharded, the candidates travel the k_persist way: every rank
// pushes header + scaled candidate row to every rank, one flag wait per pivot.
// ------------------------------------------------------------------------------------------------
// A wait of an earlier launch timed out (PSync::abort): the first look thread of the next launch
// turns that into the verdict the host is polling for, and nothing else runs.
__device__ __forceinline__ bool piter2_aborted(const PersistArgs &P)
{
    const int a = *reinterpret_cast<const volatile int *>(&P.sync->abort);
    if (a == 0) return false;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.report->status = a;
    return true;
}

template <int TR, int UNROLL, bool STREAM>
__global__ void __launch_bounds__(kPivotThreads, 2)
k_iter2(const __grid_constant__ PersistArgs P, const long long L)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    PSync *S = P.sync;
    const int G = P.look_ctas;
    {
        // look CTAs only (the tiles judge by their decision record): fin_at = f means decisions
        // 1..f-1 are pivots to apply and decision f is the verdict -- nothing left to decide
        if ((int)blockIdx.x < G) {
            if (*reinterpret_cast<const volatile long long *>(&S->carry.fin_at) != 0) return;
            if (piter2_aborted(P)) return;
        }
    }
    if ((int)blockIdx.x < G) {
        persist_look<true>(P, blockIdx.x, G, L + 1);
        return;
    }
    if (L == 0) return;                                            // nothing decided yet
    // One memory round trip before the tile's own loads: the decision record, the 16 pivot-column
    // values and the thread's slice of the pivot row are all requested together.  Plain (L1-cached)
    // loads are safe here -- everything was written by an earlier launch and L1 starts empty -- and
    // let the CTAs that share an SM share the pivot-row lines.
    __shared__ double s_col[TR];
    const int slot = (int)((L + P.slot_base) & (kRing - 1));
    const IterState *st = P.ring + slot;
    const int st_status = st->status, st_p = st->p, st_w = st->w;
    const long long st_iters = st->iters;
    const int ldv = (int)(P.ld >> 1);
    const int tiles_x = (ldv + kPivotThreads - 1) / kPivotThreads;
    const int tile = (int)blockIdx.x - G;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int r_begin = ty * TR;
    const int r_end = min(P.R_local, r_begin + TR);
    const double *col = P.colring + (int64_t)slot * P.col_stride;
    for (int t = threadIdx.x; t < TR; t += kPivotThreads)
        s_col[t] = (r_begin + t < r_end) ? col[r_begin + t] : 0.0;
    // a slot that still holds the decision of four launches ago (the solve ended, or was aborted)
    // is told apart by its pivot count
    const bool go = st_status == ST_RUNNING && st_p >= 0 && st_iters == P.iters0 + L - 1;
    if (!go) return;
    const int rel = st_p - P.row0;
    const int p_local = (rel >= 0 && rel < P.m_local) ? rel : -1;
    const double *prow = p_prow(P, slot, st_w);
    const int cv = tx * kPivotThreads + threadIdx.x;
    const double2 pr = (cv < ldv) ? *reinterpret_cast<const double2 *>(prow + 2 * cv) : make_double2(0.0, 0.0);
    __syncthreads();
    update_tile<TR, UNROLL, STREAM>(reinterpret_cast<const double2 *>(P.tab[(L - 1) & 1]),
                                    reinterpret_cast<double2 *>(P.tab[L & 1]), ldv, cv, pr, r_begin,
                                    r_end, p_local, s_col);
}

// ------------------------------------------------------------------------------------------------
// k_iter2_bulk: the same launch with the update tile staged through shared memory by the bulk
// asynchronous copy engine (TMA's 1-D form): one elected thread issues cp.async.bulk for the
// tile's row segments global -> shared (completion counted in bytes on an mbarrier), every thread
// updates its 16-byte column slice in place in shared memory, and the tile goes back with
// cp.async.bulk shared -> global.  No registers hold data in flight (3 CTAs x 64 KB per SM instead
// of 2 CTAs x 16 loads x 16 B x 256 threads).  SASS: UBLKCP (both directions), SYNCS (mbarrier).
// The A/B against the LDG/STG tile is in profiles/ (variant 30 vs 10); see DESIGN.md.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkRows = 16;
constexpr int kBulkRowBytes = kPivotThreads * 16;      // 4 KB: one row segment of a tile

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

template <bool STREAM>
__global__ void __launch_bounds__(kPivotThreads, 3)
k_iter2_bulk(const __grid_constant__ PersistArgs P, const long long L)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    PSync *S = P.sync;
    const int G = P.look_ctas;
    if ((int)blockIdx.x < G) {
        if (*reinterpret_cast<const volatile long long *>(&S->carry.fin_at) != 0) return;
        if (piter2_aborted(P)) return;
        persist_look<true>(P, blockIdx.x, G, L + 1);
        return;
    }
    if (L == 0) return;
    extern __shared__ __align__(128) unsigned char bulk_smem[];   // kBulkRows x 4 KB
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double s_col[kBulkRows];
    const int slot = (int)((L + P.slot_base) & (kRing - 1));
    const IterState *st = P.ring + slot;
    const int st_status = st->status, st_p = st->p, st_w = st->w;
    const long long st_iters = st->iters;
    const bool go = st_status == ST_RUNNING && st_p >= 0 && st_iters == P.iters0 + L - 1;
    if (!go) return;
    const int ldv = (int)(P.ld >> 1);
    const int tiles_x = (ldv + kPivotThreads - 1) / kPivotThreads;
    const int tile = (int)blockIdx.x - G;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int r_begin = ty * kBulkRows;
    const int rows = min(P.R_local, r_begin + kBulkRows) - r_begin;
    const int width = min(kPivotThreads, ldv - tx * kPivotThreads);       // double2 columns in this tile
    const uint32_t row_bytes = (uint32_t)width * 16u;                      // multiple of 128
    const double *src = P.tab[(L - 1) & 1];
    double *dst = P.tab[L & 1];
    const int64_t col0 = (int64_t)tx * kPivotThreads * 2;                  // first double column
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(row_bytes * (uint32_t)rows) : "memory");
        for (int r = 0; r < rows; ++r) {
            const double *g = src + (int64_t)(r_begin + r) * P.ld + col0;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(bulk_smem + (size_t)r * kBulkRowBytes)), "l"(g), "r"(row_bytes), "r"(bar)
                : "memory");
        }
    }
    const double *col = P.colring + (int64_t)slot * P.col_stride;
    for (int t = threadIdx.x; t < kBulkRows; t += kPivotThreads)
        s_col[t] = (t < rows) ? col[r_begin + t] : 0.0;
    const int rel = st_p - P.row0;
    const int p_local = (rel >= 0 && rel < P.m_local) ? rel : -1;
    const double *prow = p_prow(P, slot, st_w);
    const int cv = tx * kPivotThreads + threadIdx.x;
    const bool act = (int)threadIdx.x < width;
    const double2 pr = act ? *reinterpret_cast<const double2 *>(prow + 2 * cv) : make_double2(0.0, 0.0);
    __syncthreads();                                   // s_col; the barrier is initialised for everyone
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    }
    if (act) {
        double2 *tilev = reinterpret_cast<double2 *>(bulk_smem);
#pragma unroll
        for (int r = 0; r < kBulkRows; ++r) {
            if (r < rows) {
                double2 *cell = tilev + r * kPivotThreads + threadIdx.x;
                const double2 a = *cell;
                const double t = s_col[r];
                double2 o;
                o.x = __dsub_rn(a.x, __dmul_rn(t, pr.x));
                o.y = __dsub_rn(a.y, __dmul_rn(t, pr.y));
                if (r_begin + r == p_local) o = pr;
                *cell = o;
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> async proxy reads
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < rows; ++r) {
            double *g = dst + (int64_t)(r_begin + r) * P.ld + col0;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(g), "r"(smem_u32(bulk_smem + (size_t)r * kBulkRowBytes)), "r"(row_bytes)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may go once it is read
    }
    (void)STREAM;
}

} // namespace b200lp
