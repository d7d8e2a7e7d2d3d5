// solver.cu -- host side of libb200lp.so: the device-resident simplex loop and the C ABI
// declared in include/b200lp.h.
//
// Replaces the reference's n-solve-tableau (src/simplex.lisp:399-461).  The tableau lives in
// HBM for the whole solve (two ping-pong buffers).  Three packagings of the loop, chosen by shape
// (want_persist / small_ok below; measurements in profiles/r02_loop_ab_*.json):
//   k_iter2   (persist.cuh)  one launch per pivot with programmatic dependent launch: look CTAs
//             decide pivot L+1 while one CTA per 16-row tile applies pivot L.  The host enqueues
//             launches ahead (every CTA returns at once when its decision record says the solve
//             is over) and polls a Report word every `poll_interval` pivots, one batch behind the
//             enqueue front, so the GPU queue never drains.  Tableaus streamed from HBM.
//   k_persist (persist.cuh)  one cooperative kernel per b200lp_iterate call.  L2-resident tableaus.
//   k_small   (small.cuh)    the whole solve in one CTA's shared memory.  The reference's usual sizes.
// k_iter / k_look + k_update (kernels.cuh) are round 1's look role, kept as B200LP_LOOK=1 and as
// the NCCL fallback.
//
// Sharding: contiguous row blocks (b200lp_partition), objective row replicated on every shard.
// The per-pivot exchange of candidate pivot rows happens inside the look role through peer-mapped
// buffers (setup_exchange: peer access in-process, CUDA IPC handles across processes); when
// peers cannot be mapped, k_look / k_update run on two streams with an NCCL all-gather between
// them.  A shard is one GPU: either one per process (b200lp_create_sharded, torchrun style) or
// several in this process (opts.ndev > 1).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200lp.h"
#include "kernels.cuh"
#include "persist.cuh"
#include "small.cuh"
#include "nccl_dyn.h"

namespace b200lp {

// CL double-float-epsilon (SBCL): 2^-53 (1 + 2^-52); src/utils.lisp:92,107 scale it by `factor`.
static constexpr double kClEps = 0x1.0000000000001p-53;

static thread_local std::string g_last_error;
static const bool g_use_pdl = []() {
    const char *e = getenv("B200LP_PDL");       // default on; B200LP_PDL=0 serialises iterations
    return !(e && e[0] == '0');
}();
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
// One pipelined loop per GPU at a time: the lookahead CTAs of k_iter meet at grid barriers and
// must not compete for SM slots with another handle's tiles.  Different handles on the same
// device therefore serialise at b200lp_iterate granularity ("serialise rather than corrupt").
static constexpr int kMaxDeviceLocks = 64;
static std::mutex g_device_mu[kMaxDeviceLocks];

static int fail(int code, const char *what, const char *detail)
{
    g_last_error = std::string(what) + ": " + (detail ? detail : "");
    return code;
}

#define CU_TRY(expr)                                                                     \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            (void)cudaGetLastError();                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? B200LP_ERR_OUT_OF_MEMORY      \
                                                         : B200LP_ERR_CUDA,              \
                        #expr, cudaGetErrorString(e__));                                 \
        }                                                                                \
    } while (0)

#define NCCL_TRY(expr)                                                                   \
    do {                                                                                 \
        ncclResult_t r__ = (expr);                                                       \
        if (r__ != ncclSuccess)                                                          \
            return fail(B200LP_ERR_NCCL, #expr, g_nccl.GetErrorString(r__));             \
    } while (0)

#define RC_TRY(expr)                                                                     \
    do {                                                                                 \
        int rc__ = (expr);                                                               \
        if (rc__ != B200LP_OK) return rc__;                                              \
    } while (0)

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Shard {
    int device = 0;
    int rank = 0;               // global rank of this shard
    int64_t row0 = 0;           // first global constraint row owned
    int m_local = 0;            // constraint rows owned
    int R_local = 0;            // m_local + 1 (objective replica last)
    int64_t ld = 0;             // device leading dimension (multiple of 16 doubles)
    // allocation capacities (>= the current shape; a pooled handle is re-bound to other shapes)
    int64_t cap_rows = 0, cap_ld = 0;
    int cap_trace = 0;
    cudaStream_t stream = nullptr;      // uploads, step-by-step API, k_update
    cudaStream_t look_stream = nullptr; // k_look + candidate exchange (high priority)
    double *tabs[2] = {nullptr, nullptr}; // ping-pong tableau buffers; [1] allocated on first iterate
    int cur = 0;                // which buffer holds the current tableau
    double *tab = nullptr;      // == tabs[cur]
    int32_t *basis = nullptr;
    // kRing-slot decision ring of the pipelined loop (slot = iteration & 3); the step-by-step
    // API uses slot 0 through the aliases below
    double *colring = nullptr;  // kRing x R_local
    double *candring = nullptr; // kRing x (kCandHdr + ld)
    double *gathring = nullptr; // kRing x world x (kCandHdr + ld)   (exchange mode 1 only)
    IterState *ring = nullptr;  // kRing slots
    LookSync *look_sync = nullptr;
    int look_ctas = 1;
    int iter_ctas = 0;          // grid of k_iter (look_ctas + persistent update CTAs)
    double *xbuf = nullptr;     // peer-mapped exchange buffer (exchange mode 2), see struct Xchg
    Xchg xchg;                  // every rank's xbuf as mapped here
    bool ipc_opened[kMaxWorld] = {false, false, false, false, false, false, false, false};
    Report *report = nullptr;
    Report *h_report = nullptr; // pinned, 2 slots
    cudaEvent_t ev_look[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_upd[4] = {nullptr, nullptr, nullptr, nullptr};
    double *colbuf = nullptr;   // == colring slot 0
    double *cand = nullptr;     // == candring slot 0: CandHdr + ld doubles
    double *gathered = nullptr; // == gathring slot 0: world * (kCandHdr + ld) doubles
    double *colout = nullptr;   // R_local doubles, RHS gather
    DevState *st = nullptr;
    Cand *partials = nullptr;
    int2 *trace = nullptr;
    DevState *h_st = nullptr;   // pinned, 2 slots
    int2 *h_trace = nullptr;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_poll[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pivot; // pairs
    std::vector<cudaEvent_t> ev_lookt; // triples: before look, after look, after exchange
    int ratio_blocks = 0;
    // persistent cooperative loop (persist.cuh)
    double *objc = nullptr;     // ld doubles: running objective row of the look role
    double *rhsc = nullptr;     // R_local doubles: running RHS column
    PSync *psync = nullptr;
    double *rates = nullptr;    // per-CTA streaming rates of k_persist's tile role (4096 slots)
    unsigned long long *tile_prof = nullptr;   // B200LP_TILE_PROFILE only
    int coop = 0;               // cudaDevAttrCooperativeLaunch
    int sm_count = 0;
    int plook_ctas = 1;
    int last_cluster = 0;       // cluster size of the last k_persist launch (0 = no clusters)
    PersistArgs pargs;          // arguments of the current b200lp_iterate call (k_iter2)
    int iter2_ctas = 0;
};

} // namespace b200lp

using namespace b200lp;

struct b200lp_solver {
    b200lp_opts opts;
    int64_t R = 0, C = 0, m = 0;
    int is_max = 1;
    double tol = 1024.0, thr_enter = 0, thr_pivot = 0, thr_feas = 0;
    int world = 1;
    bool multiprocess = false;
    std::vector<Shard> shards;
    std::mutex mu;
    long long iters_done = 0;   // host mirror of DevState::iters
    int64_t kernel_launches = 0;
    int64_t pivot_launches_timed = 0;
    size_t ev_used = 0;
    size_t evl_used = 0;
    // how sharded candidates travel: 0 one shard, 1 NCCL all-gather between two kernels,
    // 2 peer-mapped buffers written by the look role inside k_iter
    int xmode = 0;
    unsigned long long epoch = 0;
    unsigned long long peer_timeout_ns = 120ull * 1000 * 1000 * 1000;   // B200LP_PEER_TIMEOUT_MS
    unsigned long long spin_timeout_ns = 20ull * 1000 * 1000 * 1000;    // B200LP_SPIN_TIMEOUT_MS
    unsigned long long ring_base = 0;   // decisions published so far (persistent loop's ring position)
    int last_loop = 0;                  // 1 = k_iter per pivot, 2 = k_persist
    bool cur_look2 = false;             // this call's per-pivot launches are k_iter2 (look role of persist.cuh)
};

namespace b200lp {

static void fill_thresholds(b200lp_solver *s)
{
    s->tol = s->opts.fp_tolerance_factor > 0 ? s->opts.fp_tolerance_factor : 1024.0;
    b200lp_thresholds(s->tol, &s->thr_enter, &s->thr_pivot, &s->thr_feas);
}

// Small single-GPU handles are pooled by the one-shot calls (create/destroy costs ~3 ms, a
// tiny solve ~0.3 ms); their buffers are allocated with some slack so that nearby shapes --
// branch-and-bound nodes add a row and a column per level -- reuse them.
static constexpr double kPoolMaxBytes = 64e6;
static bool poolable(const b200lp_solver *s, const Shard &sh)
{
    return s->world == 1 && 8.0 * (double)sh.ld * sh.R_local <= kPoolMaxBytes;
}

// Everything that depends only on the current shape (R_local, ld) of a shard.
static void shape_shard(Shard &sh)
{
    // enough look CTAs that each thread touches only a few cells per scan; one CTA (no grid
    // barriers) when the rows are short
    sh.look_ctas = sh.ld <= 4096 ? 1 : (int)std::min<int64_t>(kLookMaxCtas, (sh.ld + 1023) / 1024);
    sh.iter_ctas = 0;
    sh.iter2_ctas = 0;
    sh.ratio_blocks = (sh.R_local + kRatioThreads - 1) / kRatioThreads;
    sh.xchg.ld = sh.ld;
    // persistent loop: one look CTA while a row is short (no look-grid barriers at all), else
    // enough that a look thread touches only a few 16-byte units per phase
    // (measured, profiles/r02_loop_ab_*.json: 784-double rows 6.1 us per pivot with one look CTA,
    // 6.9-7.7 us with more; 3 088-double rows 11.6 us with one, 9.5 us with eight meeting at
    // global-memory barriers, 8.6 us with eight as one thread-block cluster)
    // Eight CTAs already give every look thread a single 16-byte unit per phase up to 4 096-double
    // rows and a handful beyond; more CTAs only take slots from the update tiles -- sharded, the look
    // CTAs spend most of a pivot waiting for the slowest peer's candidate.
    sh.plook_ctas = sh.ld <= 1536 ? 1 : sh.ld <= 32768 ? 8
                  : (int)std::min<int64_t>(kPLookMax, (sh.ld + 4095) / 4096);
    if (const char *e = getenv("B200LP_LOOK_CTAS")) {
        const int g = atoi(e);
        if (g >= 1 && g <= kPLookMax) sh.plook_ctas = g;
    }
}

static int ensure_trace(b200lp_solver *s, Shard &sh)
{
    const int tc = s->opts.trace_capacity;
    if (tc <= sh.cap_trace) return B200LP_OK;
    cudaFree(sh.trace);
    if (sh.h_trace) cudaFreeHost(sh.h_trace);
    sh.trace = nullptr; sh.h_trace = nullptr; sh.cap_trace = 0;
    CU_TRY(cudaMalloc(&sh.trace, sizeof(int2) * tc));
    CU_TRY(cudaMallocHost(&sh.h_trace, sizeof(int2) * tc));
    sh.cap_trace = tc;
    return B200LP_OK;
}

static int alloc_shard(b200lp_solver *s, Shard &sh)
{
    CU_TRY(cudaSetDevice(sh.device));
    int prio_lo = 0, prio_hi = 0;
    CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU_TRY(cudaStreamCreateWithPriority(&sh.stream, cudaStreamNonBlocking, prio_lo));
    CU_TRY(cudaStreamCreateWithPriority(&sh.look_stream, cudaStreamNonBlocking, prio_hi));
    sh.ld = round_up(s->C, 16);
    sh.cap_rows = sh.R_local;
    sh.cap_ld = sh.ld;
    if (poolable(s, sh)) {                                 // room for nearby shapes (pool reuse)
        sh.cap_rows = round_up(sh.R_local, 64);
        sh.cap_ld = round_up(sh.ld, 256);
    }
    const int64_t stride = kCandHdr + sh.cap_ld;
    CU_TRY(cudaMalloc(&sh.tabs[0], sizeof(double) * sh.cap_ld * sh.cap_rows));
    sh.cur = 0;
    sh.tab = sh.tabs[0];
    CU_TRY(cudaMalloc(&sh.basis, sizeof(int32_t) * sh.cap_rows));
    CU_TRY(cudaMalloc(&sh.colout, sizeof(double) * sh.cap_rows));
    CU_TRY(cudaMalloc(&sh.colring, sizeof(double) * sh.cap_rows * kRing));
    CU_TRY(cudaMalloc(&sh.candring, sizeof(double) * stride * kRing));
    CU_TRY(cudaMemsetAsync(sh.candring, 0, sizeof(double) * stride * kRing, sh.stream));
    if (s->world > 1) {
        CU_TRY(cudaMalloc(&sh.gathring, sizeof(double) * stride * s->world * kRing));
        CU_TRY(cudaMemsetAsync(sh.gathring, 0, sizeof(double) * stride * s->world * kRing, sh.stream));
        const int64_t xw = std::max(xchg_words(sh.ld), px_words(sh.ld));
        CU_TRY(cudaMalloc(&sh.xbuf, sizeof(double) * xw));
        CU_TRY(cudaMemsetAsync(sh.xbuf, 0, sizeof(double) * xw, sh.stream));
    }
    CU_TRY(cudaMalloc(&sh.objc, sizeof(double) * sh.cap_ld));
    CU_TRY(cudaMalloc(&sh.rhsc, sizeof(double) * sh.cap_rows));
    CU_TRY(cudaMalloc(&sh.psync, sizeof(PSync)));
    CU_TRY(cudaMalloc(&sh.rates, sizeof(double) * 4096));
    CU_TRY(cudaMemsetAsync(sh.rates, 0, sizeof(double) * 4096, sh.stream));   // 0 = not measured yet
    CU_TRY(cudaDeviceGetAttribute(&sh.coop, cudaDevAttrCooperativeLaunch, sh.device));
    CU_TRY(cudaDeviceGetAttribute(&sh.sm_count, cudaDevAttrMultiProcessorCount, sh.device));
    std::memset(&sh.xchg, 0, sizeof(sh.xchg));
    sh.xchg.rank = sh.rank; sh.xchg.world = s->world; sh.xchg.ld = sh.ld;
    sh.colbuf = sh.colring;
    sh.cand = sh.candring;
    sh.gathered = sh.gathring;
    CU_TRY(cudaMalloc(&sh.ring, kRing * sizeof(IterState)));
    CU_TRY(cudaMalloc(&sh.look_sync, sizeof(LookSync)));
    CU_TRY(cudaMemsetAsync(sh.look_sync, 0, sizeof(LookSync), sh.stream));
    shape_shard(sh);
    CU_TRY(cudaMalloc(&sh.report, sizeof(Report)));
    CU_TRY(cudaMallocHost(&sh.h_report, 2 * sizeof(Report)));
    for (int k = 0; k < 4; ++k) {
        CU_TRY(cudaEventCreateWithFlags(&sh.ev_look[k], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&sh.ev_upd[k], cudaEventDisableTiming));
    }
    CU_TRY(cudaMalloc(&sh.partials, sizeof(Cand) * ((sh.cap_rows + kRatioThreads - 1) / kRatioThreads)));
    CU_TRY(cudaMalloc(&sh.st, sizeof(DevState)));
    CU_TRY(cudaMemsetAsync(sh.st, 0, sizeof(DevState), sh.stream));
    CU_TRY(cudaMallocHost(&sh.h_st, 2 * sizeof(DevState)));
    RC_TRY(ensure_trace(s, sh));
    CU_TRY(cudaEventCreate(&sh.ev_begin));
    CU_TRY(cudaEventCreate(&sh.ev_end));
    CU_TRY(cudaEventCreateWithFlags(&sh.ev_poll[0], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&sh.ev_poll[1], cudaEventDisableTiming));
    CU_TRY(cudaStreamSynchronize(sh.stream));
    return B200LP_OK;
}

static void free_shard(Shard &sh)
{
    cudaSetDevice(sh.device);
    if (sh.stream) cudaStreamSynchronize(sh.stream);
    if (sh.look_stream) cudaStreamSynchronize(sh.look_stream);
    if (sh.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(sh.comm);
    cudaFree(sh.tabs[0]); cudaFree(sh.tabs[1]); cudaFree(sh.basis); cudaFree(sh.colout);
    for (int g = 0; g < kMaxWorld; ++g)
        if (sh.ipc_opened[g] && sh.xchg.peer[g]) cudaIpcCloseMemHandle(sh.xchg.peer[g]);
    cudaFree(sh.colring); cudaFree(sh.candring); cudaFree(sh.gathring); cudaFree(sh.xbuf);
    cudaFree(sh.ring); cudaFree(sh.report); cudaFree(sh.look_sync);
    cudaFree(sh.partials); cudaFree(sh.st);
    cudaFree(sh.objc); cudaFree(sh.rhsc); cudaFree(sh.psync); cudaFree(sh.rates); cudaFree(sh.tile_prof);
    cudaFree(sh.trace);
    if (sh.h_report) cudaFreeHost(sh.h_report);
    for (int k = 0; k < 4; ++k) {
        if (sh.ev_look[k]) cudaEventDestroy(sh.ev_look[k]);
        if (sh.ev_upd[k]) cudaEventDestroy(sh.ev_upd[k]);
    }
    if (sh.look_stream) cudaStreamDestroy(sh.look_stream);
    if (sh.h_st) cudaFreeHost(sh.h_st);
    if (sh.h_trace) cudaFreeHost(sh.h_trace);
    for (cudaEvent_t e : sh.ev_pivot) cudaEventDestroy(e);
    for (cudaEvent_t e : sh.ev_lookt) cudaEventDestroy(e);
    if (sh.ev_begin) cudaEventDestroy(sh.ev_begin);
    if (sh.ev_end) cudaEventDestroy(sh.ev_end);
    for (int k = 0; k < 2; ++k) if (sh.ev_poll[k]) cudaEventDestroy(sh.ev_poll[k]);
    if (sh.stream) cudaStreamDestroy(sh.stream);
    sh = Shard();
}

// Write the loop-control words; iters is kept from the host mirror.
static int push_state(b200lp_solver *s, int status, long long max_iters, int j, int forced_row)
{
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        DevState *h = &sh.h_st[0];
        std::memset(h, 0, sizeof(DevState));
        h->status = status;
        h->j = j;
        h->p = -1;
        h->winner = 0;
        h->iters = s->iters_done;
        h->max_iters = max_iters;
        h->forced_row = forced_row;
        CU_TRY(cudaMemcpyAsync(sh.st, h, sizeof(DevState), cudaMemcpyHostToDevice, sh.stream));
        // the pinned slot is reused by polls: make sure the copy has consumed it
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    return B200LP_OK;
}

static int pull_state(b200lp_solver *s, DevState *out)
{
    Shard &sh = s->shards[0];
    CU_TRY(cudaSetDevice(sh.device));
    CU_TRY(cudaMemcpyAsync(&sh.h_st[0], sh.st, sizeof(DevState), cudaMemcpyDeviceToHost, sh.stream));
    CU_TRY(cudaStreamSynchronize(sh.stream));
    *out = sh.h_st[0];
    // the other shards finish the same kernels; wait so callers may touch their memory
    for (size_t k = 1; k < s->shards.size(); ++k) {
        CU_TRY(cudaSetDevice(s->shards[k].device));
        CU_TRY(cudaStreamSynchronize(s->shards[k].stream));
    }
    return B200LP_OK;
}

// ---- kernel launches -------------------------------------------------------------------------
static void launch_enter(b200lp_solver *s, Shard &sh)
{
    const double *obj = sh.tab + (int64_t)sh.m_local * sh.ld;
    k_enter<<<1, kEnterThreads, 0, sh.stream>>>(obj, (int)(s->C - 1), s->is_max, s->thr_enter,
                                                s->opts.pivot_rule, sh.st);
    s->kernel_launches++;
}

static void launch_ratio(b200lp_solver *s, Shard &sh, bool with_trace)
{
    k_ratio<<<sh.ratio_blocks, kRatioThreads, 0, sh.stream>>>(
        sh.tab, sh.ld, sh.m_local, sh.R_local, (int)(s->C - 1), sh.basis, (int)sh.row0,
        s->thr_pivot, s->opts.pivot_rule, s->world, sh.colbuf, sh.st, sh.partials,
        reinterpret_cast<CandHdr *>(sh.cand), with_trace ? sh.trace : nullptr,
        s->opts.trace_capacity);
    s->kernel_launches++;
}

static void launch_cand(b200lp_solver *s, Shard &sh)
{
    const int blocks = (int)std::min<int64_t>((sh.ld + 255) / 256, 148);
    k_cand<<<blocks, 256, 0, sh.stream>>>(sh.tab, sh.ld, (int)s->C, (int)sh.row0, sh.st,
                                          reinterpret_cast<const CandHdr *>(sh.cand),
                                          sh.cand + kCandHdr);
    s->kernel_launches++;
}

static void launch_winner(b200lp_solver *s, Shard &sh, bool with_trace)
{
    k_winner<<<1, 32, 0, sh.stream>>>(sh.gathered, kCandHdr + sh.ld, s->world, sh.st,
                                      with_trace ? sh.trace : nullptr, s->opts.trace_capacity);
    s->kernel_launches++;
}

template <int TR, int UNROLL, int VEC, bool STREAM>
static void launch_pivot_t(b200lp_solver *s, Shard &sh)
{
    const int ldv = (int)(sh.ld / 2);
    dim3 grid((ldv + kPivotThreads * VEC - 1) / (kPivotThreads * VEC), (sh.R_local + TR - 1) / TR);
    const double *cand_base = s->world > 1 ? sh.gathered : sh.cand;
    k_pivot<TR, UNROLL, VEC, STREAM><<<grid, kPivotThreads, 0, sh.stream>>>(
        sh.tab, sh.ld, sh.m_local, sh.R_local, (int)sh.row0, sh.colbuf, cand_base,
        kCandHdr + sh.ld, sh.basis, sh.st);
}

static void launch_pivot(b200lp_solver *s, Shard &sh)
{
    int v = s->opts.pivot_variant;
    if (v == 0) {
        // Tableaus that fit the 126 MB L2 keep default caching; larger ones stream.
        const double bytes = 8.0 * (double)sh.ld * sh.R_local;
        v = bytes > 96e6 ? 1 : 2;
    }
    switch (v) {
    default:
    case 1: launch_pivot_t<64, 8, 1, true>(s, sh); break;
    case 2: launch_pivot_t<64, 8, 1, false>(s, sh); break;
    case 3: launch_pivot_t<128, 8, 1, true>(s, sh); break;
    case 4: launch_pivot_t<64, 4, 2, true>(s, sh); break;
    case 5: launch_pivot_t<32, 8, 1, true>(s, sh); break;
    case 6: launch_pivot_t<64, 16, 1, true>(s, sh); break;
    case 7: launch_pivot_t<128, 8, 2, true>(s, sh); break;
    case 8: launch_pivot_t<64, 4, 1, true>(s, sh); break;
    case 9: launch_pivot_t<128, 16, 1, false>(s, sh); break;
    }
    s->kernel_launches++;
}

// All-gather of every shard's candidate (header + scaled row) of ring slot `slot`.  The pipelined
// loop runs it on the look stream, the step-by-step API (slot 0) on the main stream.
static int exchange(b200lp_solver *s, int slot = 0, bool on_look_stream = false)
{
    if (s->world == 1) return B200LP_OK;
    const size_t bytes = sizeof(double) * (kCandHdr + s->shards[0].ld);
    if (!s->multiprocess && !s->shards[0].comm) {
        // several GPUs in this process and no NCCL communicator (peer-mapped mode): the
        // step-by-step API gathers with peer copies, ordered by one event per source shard
        for (Shard &a : s->shards) {
            CU_TRY(cudaSetDevice(a.device));
            const int64_t stride = kCandHdr + a.ld;
            for (Shard &b : s->shards)
                CU_TRY(cudaMemcpyPeerAsync(b.gathring + (slot * s->world + a.rank) * stride, b.device,
                                           a.candring + slot * stride, a.device, bytes, a.stream));
            CU_TRY(cudaEventRecord(a.ev_look[0], a.stream));
        }
        for (Shard &b : s->shards) {
            CU_TRY(cudaSetDevice(b.device));
            for (Shard &a : s->shards)
                if (&a != &b) CU_TRY(cudaStreamWaitEvent(b.stream, a.ev_look[0], 0));
        }
        return B200LP_OK;
    }
    if (s->shards.size() > 1) NCCL_TRY(g_nccl.GroupStart());
    for (Shard &sh : s->shards) {
        if (s->shards.size() > 1) cudaSetDevice(sh.device);
        const int64_t stride = kCandHdr + sh.ld;
        NCCL_TRY(g_nccl.AllGather(sh.candring + slot * stride, sh.gathring + slot * s->world * stride,
                                  bytes, ncclChar, sh.comm,
                                  on_look_stream ? sh.look_stream : sh.stream));
    }
    if (s->shards.size() > 1) NCCL_TRY(g_nccl.GroupEnd());
    return B200LP_OK;
}

// One full simplex iteration on every local shard.  `do_enter` is false for forced pivots.
static int enqueue_iteration(b200lp_solver *s, bool do_enter, bool with_trace, bool time_pivot)
{
    const bool multi = s->shards.size() > 1;
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        if (do_enter) launch_enter(s, sh);
        launch_ratio(s, sh, with_trace);
        launch_cand(s, sh);
    }
    if (s->world > 1) {
        RC_TRY(exchange(s));
        for (Shard &sh : s->shards) {
            if (multi) CU_TRY(cudaSetDevice(sh.device));
            launch_winner(s, sh, with_trace);
        }
    }
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        const bool timed = time_pivot && &sh == &s->shards[0];
        if (timed) {
            if (sh.ev_pivot.size() < s->ev_used + 2) {
                cudaEvent_t a, b;
                CU_TRY(cudaEventCreate(&a));
                CU_TRY(cudaEventCreate(&b));
                sh.ev_pivot.push_back(a);
                sh.ev_pivot.push_back(b);
            }
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used], sh.stream));
        }
        launch_pivot(s, sh);
        if (timed) {
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used + 1], sh.stream));
            s->ev_used += 2;
        }
    }
    CU_TRY(cudaGetLastError());
    return B200LP_OK;
}

// ---- pipelined loop ------------------------------------------------------------------------------
// look(k -> k+1) reads the tableau before pivot k (buffer cur + (k-1), cur itself for k = 0) and
// ring slot k; update(k) streams that buffer into the other one.
static void fill_args(b200lp_solver *s, Shard &sh, long long k, long long cap, LookArgs *a,
                      UpdateArgs *u)
{
    const int slot = (int)(k & (kRing - 1)), out = (int)((k + 1) & (kRing - 1));
    const double *src = sh.tabs[(sh.cur + (int)((k > 0 ? k - 1 : 0) & 1)) & 1];
    const int64_t stride = kCandHdr + sh.ld;
    a->src = src; a->ld = sh.ld;
    a->C = (int)s->C; a->m_local = sh.m_local; a->R_local = sh.R_local; a->row0 = (int)sh.row0;
    a->world = s->world; a->rank = sh.rank; a->is_max = s->is_max; a->rule = s->opts.pivot_rule;
    a->slot_in = slot; a->slot_out = out;
    a->thr_enter = s->thr_enter; a->thr_pivot = s->thr_pivot;
    a->max_iters = cap;
    a->ring = sh.ring; a->colring = sh.colring; a->col_stride = sh.R_local;
    a->candring = sh.candring; a->gathring = sh.gathring; a->cand_stride = stride;
    a->xchg = sh.xchg; a->xchg.epoch = s->epoch;
    a->mode = s->xmode;
    a->basis = sh.basis; a->report = sh.report;
    a->trace = sh.trace; a->trace_cap = s->opts.trace_capacity;
    a->sync = sh.look_sync;
    a->timeout_ns = s->peer_timeout_ns;
    u->src = src;
    u->dst = sh.tabs[(sh.cur + (int)(k & 1)) & 1];
    u->ld = sh.ld; u->m_local = sh.m_local; u->R_local = sh.R_local; u->row0 = (int)sh.row0;
    u->world = s->world; u->mode = s->xmode; u->slot = slot;
    u->ring = sh.ring; u->colring = sh.colring; u->col_stride = sh.R_local;
    u->candring = sh.candring; u->gathring = sh.gathring; u->cand_stride = stride;
    u->xrow = sh.xbuf ? sh.xbuf + xchg_row_off(0, sh.ld) : nullptr;
}

static int pick_variant(const b200lp_solver *s, const Shard &sh)
{
    int v = s->opts.pivot_variant;
    if (v == 0) {
        // Two ping-pong buffers that fit the 126 MB L2 keep default caching; larger tableaus
        // stream in 16-row tiles with all 16 loads of a thread in flight before its first store
        // (best on B200 at every size > L2: least tail, profiles/r01_variant_sweep_*.json).
        const double bytes = 8.0 * (double)sh.ld * sh.R_local;
        v = bytes > 48e6 ? 10 : 2;
    }
    return v;
}

// variant -> (TR, UNROLL, STREAM); 4 and 7 are kept as aliases of their one-vector forms
#define B200LP_VARIANTS(X) \
    switch (v) {                       \
    default:                           \
    case 1: X(64, 8, true); break;     \
    case 2: X(64, 8, false); break;    \
    case 3: X(128, 8, true); break;    \
    case 4: X(64, 4, true); break;     \
    case 5: X(32, 8, true); break;     \
    case 6: X(64, 16, true); break;    \
    case 7: X(128, 8, true); break;    \
    case 8: X(64, 4, true); break;     \
    case 9: X(128, 16, false); break;  \
    case 10: X(16, 16, true); break;   \
    case 11: X(32, 16, true); break;   \
    case 12: X(32, 8, false); break;   \
    case 13: X(16, 8, true); break;    \
    }

template <int TR, int UNROLL, bool STREAM>
static cudaError_t launch_iter_t(Shard &sh, const LookArgs &a, const UpdateArgs &u)
{
    if (sh.iter_ctas == 0) {
        const int64_t ldv = sh.ld / 2;
        const int64_t ntiles = ((ldv + kPivotThreads - 1) / kPivotThreads) * ((sh.R_local + TR - 1) / TR);
        sh.iter_ctas = sh.look_ctas + (int)ntiles;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sh.iter_ctas);
    cfg.blockDim = dim3(kPivotThreads);
    cfg.stream = sh.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_iter<TR, UNROLL, STREAM>, a, u, sh.look_ctas);
}

static int launch_iter(b200lp_solver *s, Shard &sh, long long k, long long cap)
{
    LookArgs a;
    UpdateArgs u;
    fill_args(s, sh, k, cap, &a, &u);
    const int v = pick_variant(s, sh);
#define X(TR, UN, ST) CU_TRY((launch_iter_t<TR, UN, ST>(sh, a, u)))
    B200LP_VARIANTS(X)
#undef X
    s->kernel_launches++;
    return B200LP_OK;
}

static void launch_look(b200lp_solver *s, Shard &sh, long long k, long long cap)
{
    LookArgs a;
    UpdateArgs u;
    fill_args(s, sh, k, cap, &a, &u);
    k_look<<<sh.look_ctas, kLookThreads, 0, sh.look_stream>>>(a);
    s->kernel_launches++;
}

template <int TR, int UNROLL, bool STREAM>
static void launch_update_t(Shard &sh, const UpdateArgs &u)
{
    const int ldv = (int)(sh.ld / 2);
    dim3 grid((ldv + kPivotThreads - 1) / kPivotThreads, (sh.R_local + TR - 1) / TR);
    k_update<TR, UNROLL, STREAM><<<grid, kPivotThreads, 0, sh.stream>>>(u);
}

static void launch_update(b200lp_solver *s, Shard &sh, long long k)
{
    LookArgs a;
    UpdateArgs u;
    fill_args(s, sh, k, 0, &a, &u);
    const int v = pick_variant(s, sh);
#define X(TR, UN, ST) launch_update_t<TR, UN, ST>(sh, u)
    B200LP_VARIANTS(X)
#undef X
    s->kernel_launches++;
}

static int launch_iter2(b200lp_solver *s, Shard &sh, long long k);   // k_iter2, defined with the persistent loop below

// One fused iteration kernel on every local shard: update(k) || look(k -> k+1) (+ in-kernel exchange).
static int enqueue_iter(b200lp_solver *s, long long k, long long cap, bool time_pivot)
{
    const bool multi = s->shards.size() > 1;
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        const bool timed = time_pivot && &sh == &s->shards[0];
        if (timed) {
            if (sh.ev_pivot.size() < s->ev_used + 2) {
                cudaEvent_t a, b;
                CU_TRY(cudaEventCreate(&a));
                CU_TRY(cudaEventCreate(&b));
                sh.ev_pivot.push_back(a);
                sh.ev_pivot.push_back(b);
            }
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used], sh.stream));
        }
        if (s->cur_look2) RC_TRY(launch_iter2(s, sh, k));
        else RC_TRY(launch_iter(s, sh, k, cap));
        if (timed) {
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used + 1], sh.stream));
            s->ev_used += 2;
        }
    }
    return B200LP_OK;
}

// look(k -> k+1) (+ exchange of its candidates) on every local shard; it may start once
// update(k-1) has produced the tableau it reads.
static int enqueue_look(b200lp_solver *s, long long k, long long cap, bool timed = false)
{
    const bool multi = s->shards.size() > 1;
    Shard &s0 = s->shards[0];
    if (timed) {
        while (s0.ev_lookt.size() < s->evl_used + 3) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreate(&e));
            s0.ev_lookt.push_back(e);
        }
    }
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        if (k > 1) CU_TRY(cudaStreamWaitEvent(sh.look_stream, sh.ev_upd[(k - 1) & 3], 0));
        if (timed && &sh == &s0) CU_TRY(cudaEventRecord(s0.ev_lookt[s->evl_used], sh.look_stream));
        launch_look(s, sh, k, cap);
        if (timed && &sh == &s0) CU_TRY(cudaEventRecord(s0.ev_lookt[s->evl_used + 1], sh.look_stream));
    }
    RC_TRY(exchange(s, (int)((k + 1) & (kRing - 1)), true));
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        if (timed && &sh == &s0) {
            CU_TRY(cudaEventRecord(s0.ev_lookt[s->evl_used + 2], sh.look_stream));
            s->evl_used += 3;
        }
        CU_TRY(cudaEventRecord(sh.ev_look[(k + 1) & 3], sh.look_stream));
    }
    return B200LP_OK;
}

// update(k) on every local shard, after look(k-1 -> k) (+ exchange) has decided it.
static int enqueue_update(b200lp_solver *s, long long k, bool time_pivot)
{
    const bool multi = s->shards.size() > 1;
    for (Shard &sh : s->shards) {
        if (multi) CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamWaitEvent(sh.stream, sh.ev_look[k & 3], 0));
        const bool timed = time_pivot && &sh == &s->shards[0];
        if (timed) {
            if (sh.ev_pivot.size() < s->ev_used + 2) {
                cudaEvent_t a, b;
                CU_TRY(cudaEventCreate(&a));
                CU_TRY(cudaEventCreate(&b));
                sh.ev_pivot.push_back(a);
                sh.ev_pivot.push_back(b);
            }
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used], sh.stream));
        }
        launch_update(s, sh, k);
        if (timed) {
            CU_TRY(cudaEventRecord(sh.ev_pivot[s->ev_used + 1], sh.stream));
            s->ev_used += 2;
        }
        CU_TRY(cudaEventRecord(sh.ev_upd[k & 3], sh.stream));
    }
    return B200LP_OK;
}

static int default_poll_interval(const b200lp_solver *s)
{
    if (s->opts.poll_interval > 0) return s->opts.poll_interval;
    return 32;
}

// ---- persistent cooperative loop (persist.cuh) ---------------------------------------------------
// variant -> (UNROLL, STREAM) of k_persist's tile role
#define B200LP_PVARIANTS(X) \
    switch (v) {                   \
    default:                       \
    case 20: X(16, true); break;   \
    case 21: X(16, false); break;  \
    case 22: X(8, true); break;    \
    case 23: X(8, false); break;   \
    }

static int pick_pvariant(const b200lp_solver *s, const Shard &sh)
{
    const int v = s->opts.pivot_variant;
    if (v >= 20 && v <= 23) return v;
    // both ping-pong buffers L2 resident: default caching; larger tableaus stream (evict-first)
    const double bytes = 8.0 * (double)sh.ld * sh.R_local;
    return bytes > 48e6 ? 20 : 21;
}

static int loop_env()
{
    const char *e = getenv("B200LP_LOOP");       // iter | persist; default: persist where possible
    if (!e) return 0;
    if (std::strcmp(e, "iter") == 0) return 1;
    if (std::strcmp(e, "persist") == 0) return 2;
    return 0;
}

// Which packaging of the loop runs (measured on B200, profiles/r02_loop_ab_*.json):
//   * tableaus whose two ping-pong buffers stay in the 126 MB L2 are bound by the decision chain
//     and the launch gap: the persistent cooperative kernel wins (config 2: 10.2 vs 13.4 us per
//     pivot; m=256, n=512: 6.1 vs 10.8 us);
//   * anything streamed from HBM is bound by the rank-1 update, and there the hardware CTA
//     scheduler of a per-pivot launch balances the SMs better than any fixed ownership of tiles
//     (config 3: 468 vs 545 us; its 1 025-row 8-GPU shard: 62.3 vs 69.1 us): k_iter.
// B200LP_LOOP=iter|persist and explicit tile variants (1..13 k_iter, 20..23 k_persist) override.
// The NCCL fallback (xmode 1) always uses the per-pivot loop.
static constexpr double kPersistMaxBytes = 64e6;
static bool want_persist(const b200lp_solver *s)
{
    if (s->xmode == 1 || loop_env() == 1) return false;
    for (const Shard &sh : s->shards)
        if (!sh.coop) return false;
    const int v = s->opts.pivot_variant;
    if (v == 30) return false;                      // k_iter2_bulk: a per-pivot launch
    if (v >= 20 || loop_env() == 2) return true;
    if (v >= 1) return false;
    for (const Shard &sh : s->shards)
        if (8.0 * (double)sh.ld * sh.R_local > kPersistMaxBytes) return false;
    return true;
}

// phase sums are SM cycles; the (%globaltimer, clock64) pairs of the call convert them
static void fill_look_telemetry(const PSync &ps, b200lp_result *out)
{
    double ns_per_cycle = 0.52;                            // ~1.92 GHz when the call was too short to tell
    if (ps.clk1 > ps.clk0 && ps.gt1 > ps.gt0 && ps.gt1 - ps.gt0 > 20000)
        ns_per_cycle = (double)(ps.gt1 - ps.gt0) / (double)(ps.clk1 - ps.clk0);
    const double cyc_ms = ns_per_cycle * 1e-6;
    out->ms_look_kernel = (double)(ps.ns_wait_done + ps.ns_a + ps.ns_b1 + ps.ns_xwait + ps.ns_b2) * cyc_ms;
    out->look_kernel_launches = (int64_t)ps.look_count;
    out->ms_look_wait = (double)ps.ns_wait_done * cyc_ms;
    out->ms_look_ratio = (double)ps.ns_a * cyc_ms;
    out->ms_look_push = (double)ps.ns_b1 * cyc_ms;
    out->ms_look_peer_wait = (double)ps.ns_xwait * cyc_ms;
    out->ms_look_row = (double)ps.ns_b2 * cyc_ms;
    out->sm_clock_mhz = 1e3 / ns_per_cycle;
    for (int q = 0; q < 8; ++q) out->ms_look_dbg[q] = (double)ps.dbg[q] * cyc_ms;
}

static void fill_pargs(b200lp_solver *s, Shard &sh, long long start_iters, long long cap, PersistArgs *pa)
{
    PersistArgs &a = *pa;
    std::memset(&a, 0, sizeof(a));
    a.tab[0] = sh.tabs[sh.cur]; a.tab[1] = sh.tabs[sh.cur ^ 1];
    a.ld = sh.ld;
    a.C = (int)s->C; a.m_local = sh.m_local; a.R_local = sh.R_local; a.row0 = (int)sh.row0;
    a.world = s->world; a.rank = sh.rank; a.is_max = s->is_max; a.rule = s->opts.pivot_rule;
    a.mode = s->xmode;
    a.thr_enter = s->thr_enter; a.thr_pivot = s->thr_pivot;
    a.iters0 = start_iters; a.max_iters = cap;
    a.ring = sh.ring; a.colring = sh.colring; a.col_stride = sh.R_local;
    a.prowring = sh.candring + kCandHdr; a.prow_stride = kCandHdr + sh.ld;
    a.objc = sh.objc; a.rhsc = sh.rhsc;
    a.xchg = sh.xchg; a.xchg.epoch = s->epoch;
    a.basis = sh.basis; a.report = sh.report;
    a.trace = sh.trace; a.trace_cap = s->opts.trace_capacity;
    a.sync = sh.psync;
    a.timeout_ns = s->world > 1 ? std::max(s->peer_timeout_ns, s->spin_timeout_ns) : s->spin_timeout_ns;
    a.look_ctas = sh.plook_ctas;
    a.slot_base = (int)(s->ring_base & (kRing - 1));
    a.rates = getenv("B200LP_NO_BALANCE") ? nullptr : sh.rates;
}

// k_iter2: per-pivot launches with the two-phase look role of persist.cuh.  B200LP_LOOK=1 keeps
// the round-1 look role (k_iter).
static bool want_look2(const b200lp_solver *s)
{
    if (s->xmode == 1) return false;
    const char *e = getenv("B200LP_LOOK");
    return !(e && e[0] == '1');
}

template <int TR, int UNROLL, bool STREAM>
static cudaError_t launch_iter2_t(Shard &sh, long long k)
{
    if (sh.iter2_ctas == 0) {
        const int64_t ldv = sh.ld / 2;
        const int64_t ntiles = ((ldv + kPivotThreads - 1) / kPivotThreads) * ((sh.R_local + TR - 1) / TR);
        sh.iter2_ctas = sh.pargs.look_ctas + (int)ntiles;
    }
    cudaLaunchConfig_t cfg = {};
    // launch 0 only decides pivot 1: just the look CTAs, not 25 000 tile CTAs that find nothing to do
    cfg.gridDim = dim3((unsigned)(k == 0 ? sh.pargs.look_ctas : sh.iter2_ctas));
    cfg.blockDim = dim3(kPivotThreads);
    cfg.stream = sh.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_iter2<TR, UNROLL, STREAM>, sh.pargs, k);
}

// variant 30: the update tile staged by the bulk asynchronous copy engine (k_iter2_bulk)
static int launch_iter2_bulk(b200lp_solver *s, Shard &sh, long long k)
{
    static bool attr_set = false;
    const int smem = kBulkRows * kBulkRowBytes;
    if (!attr_set) {
        CU_TRY(cudaFuncSetAttribute(k_iter2_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    if (sh.iter2_ctas == 0) {
        const int64_t ldv = sh.ld / 2;
        const int64_t ntiles = ((ldv + kPivotThreads - 1) / kPivotThreads) * ((sh.R_local + kBulkRows - 1) / kBulkRows);
        sh.iter2_ctas = sh.pargs.look_ctas + (int)ntiles;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(k == 0 ? sh.pargs.look_ctas : sh.iter2_ctas));
    cfg.blockDim = dim3(kPivotThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = sh.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU_TRY(cudaLaunchKernelEx(&cfg, k_iter2_bulk<true>, sh.pargs, k));
    s->kernel_launches++;
    return B200LP_OK;
}

static int launch_iter2(b200lp_solver *s, Shard &sh, long long k)
{
    if (s->opts.pivot_variant == 30) return launch_iter2_bulk(s, sh, k);
    const int v = pick_variant(s, sh);
#define X(TR, UN, ST) CU_TRY((launch_iter2_t<TR, UN, ST>(sh, k)))
    B200LP_VARIANTS(X)
#undef X
    s->kernel_launches++;
    return B200LP_OK;
}

// k_persist with the look grid as one thread-block cluster (cooperative + cluster launch).  Any
// refusal by the runtime (occupancy query, launch) falls back to the plain cooperative launch.
template <int UNROLL, bool STREAM>
static cudaError_t launch_persist_cluster_t(Shard &sh, PersistArgs &a, bool *launched)
{
    *launched = false;
    const int G = a.look_ctas;
    if (G < 2 || G > 8 || (G & (G - 1)) != 0) return cudaSuccess;          // portable cluster sizes
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kPivotThreads);
    cfg.gridDim = dim3((unsigned)(G * sh.sm_count));                        // placeholder for the query
    cfg.stream = sh.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_persist_cl<UNROLL, STREAM>, &cfg) != cudaSuccess ||
        nclusters < 2) {
        (void)cudaGetLastError();
        return cudaSuccess;
    }
    cfg.gridDim = dim3((unsigned)(nclusters * G));
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_persist_cl<UNROLL, STREAM>, a);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return cudaSuccess; }
    *launched = true;
    return cudaSuccess;
}

static bool cluster_look_enabled()
{
    const char *e = getenv("B200LP_CLUSTER");            // B200LP_CLUSTER=0: global-memory barriers only
    return !(e && e[0] == '0');
}

template <int UNROLL, bool STREAM>
static cudaError_t launch_persist_t(Shard &sh, PersistArgs &a)
{
    if (a.mode == 0 && a.look_ctas > 1 && cluster_look_enabled()) {
        bool launched = false;
        cudaError_t e = launch_persist_cluster_t<UNROLL, STREAM>(sh, a, &launched);
        if (e != cudaSuccess) return e;
        sh.last_cluster = launched ? a.look_ctas : 0;
        if (launched) return cudaSuccess;
    } else {
        sh.last_cluster = 0;
    }
    static int occ = 0;                                   // CTAs per SM, same on every device here
    if (occ == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, k_persist<UNROLL, STREAM>, kPivotThreads, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
    }
    int grid = occ * sh.sm_count;
    if (a.look_ctas >= grid) a.look_ctas = 1;
    void *params[] = {&a};
    return cudaLaunchCooperativeKernel((const void *)k_persist<UNROLL, STREAM>, dim3((unsigned)grid),
                                       dim3(kPivotThreads), params, 0, sh.stream);
}

static int iterate_persist(b200lp_solver *s, int64_t limit, b200lp_result *out, int32_t *trace_j,
                           int32_t *trace_r)
{
    const double t0 = now_ms();
    if (limit <= 0) limit = s->opts.max_iters;
    const long long start_iters = s->iters_done;
    const long long cap = limit > 0 ? start_iters + limit : 0;
    const int64_t launches0 = s->kernel_launches;
    s->epoch += 1ull << 40;                                // fresh flag values for this call
    s->last_loop = 2;
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        if (!sh.tabs[1]) CU_TRY(cudaMalloc(&sh.tabs[1], sizeof(double) * sh.cap_ld * sh.cap_rows));
        Report rep;
        rep.status = ST_RUNNING; rep.pad = 0; rep.iters = start_iters;
        CU_TRY(cudaMemcpyAsync(sh.report, &rep, sizeof(rep), cudaMemcpyHostToDevice, sh.stream));
        CU_TRY(cudaMemsetAsync(sh.psync, 0, sizeof(PSync), sh.stream));
    }
    Shard &s0 = s->shards[0];
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventRecord(s0.ev_begin, s0.stream));
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        PersistArgs a;
        fill_pargs(s, sh, start_iters, cap, &a);
        if (getenv("B200LP_TILE_PROFILE")) {
            if (!sh.tile_prof) CU_TRY(cudaMalloc(&sh.tile_prof, sizeof(unsigned long long) * 2 * 4096));
            CU_TRY(cudaMemsetAsync(sh.tile_prof, 0, sizeof(unsigned long long) * 2 * 4096, sh.stream));
            a.tile_prof = sh.tile_prof;
        }
        const int v = pick_pvariant(s, sh);
#define X(UN, ST) CU_TRY((launch_persist_t<UN, ST>(sh, a)))
        B200LP_PVARIANTS(X)
#undef X
        s->kernel_launches++;
    }
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventRecord(s0.ev_end, s0.stream));
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    CU_TRY(cudaSetDevice(s0.device));
    Report last;
    PSync ps;
    std::memset(&ps, 0, sizeof(ps));
    CU_TRY(cudaMemcpy(&last, s0.report, sizeof(last), cudaMemcpyDeviceToHost));
    int aborted = 0;
    for (Shard &sh : s->shards) {                          // any shard's abort fails the call
        PSync q;
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaMemcpy(&q, sh.psync, sizeof(q), cudaMemcpyDeviceToHost));
        if (&sh == &s0) ps = q;
        if (q.abort && !aborted) aborted = q.abort;
    }
    CU_TRY(cudaSetDevice(s0.device));
    if (const char *path = getenv("B200LP_TILE_PROFILE")) {
        if (s0.tile_prof) {                                // dev aid: busy ns and SM id per CTA of the last call
            std::vector<unsigned long long> tp(2 * 4096);
            CU_TRY(cudaMemcpy(tp.data(), s0.tile_prof, sizeof(unsigned long long) * tp.size(), cudaMemcpyDeviceToHost));
            if (FILE *f = std::fopen(path, "w")) {
                for (int c = 0; c < 4096; ++c)
                    if (tp[2 * c]) std::fprintf(f, "%d,%llu,%llu\n", c, tp[2 * c + 1], tp[2 * c]);
                std::fclose(f);
            }
        }
    }
    if (aborted == ST_PEER_TIMEOUT)
        return fail(B200LP_ERR_PEER_TIMEOUT, "iterate", "a peer GPU's candidate never arrived");
    if (aborted)
        return fail(B200LP_ERR_INTERNAL, "iterate", "a wait inside the persistent loop timed out");
    if (last.status == ST_RUNNING)
        return fail(B200LP_ERR_INTERNAL, "iterate", "device loop ended without a verdict");
    const long long done = last.iters - start_iters;
    s->iters_done = last.iters;
    s->ring_base += (unsigned long long)done + 1ull;
    for (Shard &sh : s->shards) {                          // pivots ping-pong between the buffers
        sh.cur = (sh.cur + (int)(done & 1)) & 1;
        sh.tab = sh.tabs[sh.cur];
    }
    const int status = last.status;
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->status = status;
        out->n_devices = s->world;
        out->exchange_mode = s->xmode;
        out->loop_mode = 2;
        out->look_ctas = s0.plook_ctas;
        out->look_cluster = s0.last_cluster;
        out->iterations = done;
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, s0.ev_begin, s0.ev_end));
        out->ms_solve = ms;
        fill_look_telemetry(ps, out);
        out->kernel_launches = s->kernel_launches - launches0;
        out->bytes_per_pivot = 16 * (int64_t)s0.R_local * s->C;
        double obj = 0.0;
        CU_TRY(cudaMemcpy(&obj, s0.tab + (int64_t)s0.m_local * s0.ld + (s->C - 1), sizeof(double),
                          cudaMemcpyDeviceToHost));
        out->objective = obj;
        out->d2h_bytes += sizeof(double);
        int tl = 0;
        if (s->opts.trace_capacity > 0 && s0.trace) {
            tl = (int)std::min<long long>(s->iters_done, s->opts.trace_capacity);
            if (tl > 0 && (trace_j || trace_r)) {
                CU_TRY(cudaMemcpy(s0.h_trace, s0.trace, sizeof(int2) * tl, cudaMemcpyDeviceToHost));
                for (int q = 0; q < tl; ++q) {
                    if (trace_j) trace_j[q] = s0.h_trace[q].x;
                    if (trace_r) trace_r[q] = s0.h_trace[q].y;
                }
                out->d2h_bytes += sizeof(int2) * tl;
            }
        }
        out->trace_len = tl;
        out->ms_total = now_ms() - t0;
    }
    return status;
}

// n-solve-tableau's loop, src/simplex.lisp:455-460, for at most `limit` more pivots.
// Iteration k = { update(k) on the main stream || look(k -> k+1) on the look stream }; the host
// enqueues a batch ahead and polls the device Report one batch behind, so neither stream drains.
static int iterate_locked(b200lp_solver *s, int64_t limit, b200lp_result *out, int32_t *trace_j,
                          int32_t *trace_r)
{
    if (want_persist(s)) return iterate_persist(s, limit, out, trace_j, trace_r);
    const double t0 = now_ms();
    if (limit <= 0) limit = s->opts.max_iters;
    const long long start_iters = s->iters_done;
    const long long cap = limit > 0 ? start_iters + limit : 0;
    const int64_t launches0 = s->kernel_launches;
    s->last_loop = 1;
    s->ev_used = 0;
    s->evl_used = 0;
    const bool fused = s->xmode != 1;
    s->cur_look2 = fused && want_look2(s);
    s->epoch += 1ull << 40;                                // fresh sequence numbers for this call
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        if (!sh.tabs[1]) CU_TRY(cudaMalloc(&sh.tabs[1], sizeof(double) * sh.cap_ld * sh.cap_rows));
        cudaStream_t st0 = fused ? sh.stream : sh.look_stream;
        if (s->cur_look2) {
            fill_pargs(s, sh, start_iters, cap, &sh.pargs);
            sh.iter2_ctas = 0;
            CU_TRY(cudaMemsetAsync(sh.psync, 0, sizeof(PSync), st0));
        }
        IterState init;
        std::memset(&init, 0, sizeof(init));
        init.status = ST_START; init.j = -1; init.p = -1; init.iters = start_iters;
        Report rep;
        rep.status = ST_RUNNING; rep.pad = 0; rep.iters = start_iters;
        CU_TRY(cudaMemcpyAsync(sh.ring, &init, sizeof(init), cudaMemcpyHostToDevice, st0));
        CU_TRY(cudaMemcpyAsync(sh.report, &rep, sizeof(rep), cudaMemcpyHostToDevice, st0));
        CU_TRY(cudaMemsetAsync(sh.look_sync, 0, sizeof(LookSync), st0));
        CU_TRY(cudaStreamSynchronize(sh.look_stream));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    const bool time_pivot = s->opts.time_kernels != 0;
    const int batch = default_poll_interval(s);
    Shard &s0 = s->shards[0];
    cudaStream_t poll_stream = fused ? s0.stream : s0.look_stream;
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventRecord(s0.ev_begin, s0.stream));
    // iteration 0 decides iteration 1 (fused: its update role finds nothing pending)
    if (fused) RC_TRY(enqueue_iter(s, 0, cap, false));
    else RC_TRY(enqueue_look(s, 0, cap));

    Report last;
    last.status = ST_RUNNING; last.pad = 0; last.iters = start_iters;
    int slot = 0;
    bool pending[2] = {false, false};
    long long k = 0;                                       // iterations enqueued
    int batches = 0;
    const bool opts_batch = s->opts.poll_interval > 0;     // an explicit interval is taken literally
    bool all_enqueued = false;
    for (;;) {
        if (!all_enqueued) {
            // With a cap, `limit` iterations decide everything: look(limit -> limit+1) tells
            // "optimal" from "iteration limit".  Without a cap keep the queue one poll ahead.
            // ramp 4, 8, 16, ... up to the poll interval: a solve of a few pivots should not
            // pay for dozens of launches that find the solve already over
            long long nb = opts_batch ? batch : std::min<long long>(batch, 4ll << std::min(batches, 8));
            ++batches;
            if (limit > 0) nb = std::min<long long>(nb, limit - k);
            for (long long b = 0; b < nb; ++b) {
                ++k;
                if (fused) {
                    RC_TRY(enqueue_iter(s, k, cap, time_pivot && s->ev_used < 2 * 8192));
                } else {
                    RC_TRY(enqueue_update(s, k, time_pivot && s->ev_used < 2 * 8192));
                    RC_TRY(enqueue_look(s, k, cap, time_pivot && s->evl_used < 3 * 8192));
                }
            }
            if (s->shards.size() > 1) CU_TRY(cudaSetDevice(s0.device));
            CU_TRY(cudaMemcpyAsync(&s0.h_report[slot], s0.report, sizeof(Report),
                                   cudaMemcpyDeviceToHost, poll_stream));
            CU_TRY(cudaEventRecord(s0.ev_poll[slot], poll_stream));
            pending[slot] = true;
            if (limit > 0 && k >= limit) all_enqueued = true;
        }
        const int other = slot ^ 1;
        if (pending[other]) {
            CU_TRY(cudaEventSynchronize(s0.ev_poll[other]));
            pending[other] = false;
            last = s0.h_report[other];
            if (last.status != ST_RUNNING) break;
        }
        if (all_enqueued) break;   // the newest poll (drained below) is final
        slot = other;
    }
    // drain: the newest poll holds the final state
    for (int q = 0; q < 2; ++q) {
        if (pending[q]) {
            CU_TRY(cudaEventSynchronize(s0.ev_poll[q]));
            if (last.status == ST_RUNNING) last = s0.h_report[q];
        }
    }
    CU_TRY(cudaGetLastError());
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.look_stream));
    }
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventRecord(s0.ev_end, s0.stream));
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    CU_TRY(cudaSetDevice(s0.device));
    if (last.status == ST_RUNNING)
        return fail(B200LP_ERR_INTERNAL, "iterate", "device loop ended without a verdict");
    if (last.status == ST_PEER_TIMEOUT)
        return fail(B200LP_ERR_PEER_TIMEOUT, "iterate", "a peer GPU's candidate never arrived");
    if (last.status == ST_BARRIER_TIMEOUT)
        return fail(B200LP_ERR_INTERNAL, "iterate", "a barrier of the look CTAs timed out");
    if (s->cur_look2) {                                    // k_iter2: an aborted wait leaves no verdict
        for (Shard &sh : s->shards) {
            PSync q;
            CU_TRY(cudaSetDevice(sh.device));
            CU_TRY(cudaMemcpy(&q, sh.psync, sizeof(q), cudaMemcpyDeviceToHost));
            if (q.abort == ST_PEER_TIMEOUT)
                return fail(B200LP_ERR_PEER_TIMEOUT, "iterate", "a peer GPU's candidate never arrived");
            if (q.abort)
                return fail(B200LP_ERR_INTERNAL, "iterate", "a wait of the look CTAs timed out");
        }
        CU_TRY(cudaSetDevice(s0.device));
    }
    const long long done = last.iters - start_iters;
    s->iters_done = last.iters;
    if (s->cur_look2) s->ring_base += (unsigned long long)done + 1ull;
    for (Shard &sh : s->shards) {                          // pivots ping-pong between the buffers
        sh.cur = (sh.cur + (int)(done & 1)) & 1;
        sh.tab = sh.tabs[sh.cur];
    }
    const int status = last.status;

    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->status = status;
        out->n_devices = s->world;
        out->exchange_mode = s->xmode;
        out->loop_mode = 1;
        out->look_ctas = s->cur_look2 ? s0.plook_ctas : s0.look_ctas;
        out->iterations = done;
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, s0.ev_begin, s0.ev_end));
        out->ms_solve = ms;
        double pk = 0.0;
        // only launches that did real work: the first `iterations` timed pairs
        const size_t real = (size_t)std::min<long long>(out->iterations, (long long)(s->ev_used / 2));
        for (size_t q = 0; q < real; ++q) {
            float e = 0.f;
            CU_TRY(cudaEventElapsedTime(&e, s0.ev_pivot[2 * q], s0.ev_pivot[2 * q + 1]));
            pk += e;
        }
        out->ms_pivot_kernel = pk;
        out->pivot_kernel_launches = (int64_t)real;
        const size_t real_l = (size_t)std::min<long long>(std::max<long long>(out->iterations - 1, 0),
                                                          (long long)(s->evl_used / 3));
        double lk = 0.0, xk = 0.0;
        for (size_t q = 0; q < real_l; ++q) {
            float e = 0.f;
            CU_TRY(cudaEventElapsedTime(&e, s0.ev_lookt[3 * q], s0.ev_lookt[3 * q + 1]));
            lk += e;
            CU_TRY(cudaEventElapsedTime(&e, s0.ev_lookt[3 * q + 1], s0.ev_lookt[3 * q + 2]));
            xk += e;
        }
        out->ms_look_kernel = lk;
        out->ms_exchange = xk;
        out->look_kernel_launches = (int64_t)real_l;
        if (fused && !s->cur_look2) {                      // timed on the device by the look role
            LookSync ls;
            CU_TRY(cudaMemcpy(&ls, s0.look_sync, sizeof(ls), cudaMemcpyDeviceToHost));
            out->ms_look_kernel = (double)ls.look_ns * 1e-6;
            out->look_kernel_launches = ls.look_count;
        }
        if (s->cur_look2) {
            PSync ps;
            CU_TRY(cudaMemcpy(&ps, s0.psync, sizeof(ps), cudaMemcpyDeviceToHost));
            fill_look_telemetry(ps, out);
        }
        out->kernel_launches = s->kernel_launches - launches0;
        out->bytes_per_pivot = 16 * (int64_t)s0.R_local * s->C;
        double obj = 0.0;
        CU_TRY(cudaMemcpy(&obj, s0.tab + (int64_t)s0.m_local * s0.ld + (s->C - 1), sizeof(double),
                          cudaMemcpyDeviceToHost));
        out->objective = obj;
        out->d2h_bytes += sizeof(double);
        int tl = 0;
        if (s->opts.trace_capacity > 0 && s0.trace) {
            tl = (int)std::min<long long>(s->iters_done, s->opts.trace_capacity);
            if (tl > 0 && (trace_j || trace_r)) {
                CU_TRY(cudaMemcpy(s0.h_trace, s0.trace, sizeof(int2) * tl, cudaMemcpyDeviceToHost));
                for (int q = 0; q < tl; ++q) {
                    if (trace_j) trace_j[q] = s0.h_trace[q].x;
                    if (trace_r) trace_r[q] = s0.h_trace[q].y;
                }
                out->d2h_bytes += sizeof(int2) * tl;
            }
        }
        out->trace_len = tl;
        out->ms_total = now_ms() - t0;
    }
    return status;
}

static int upload_locked(b200lp_solver *s, const double *tab, int64_t ld, const int32_t *basis)
{
    if (!tab || ld < s->C) return fail(B200LP_ERR_INVALID_ARG, "upload", "ld < C or null tableau");
    const size_t wbytes = sizeof(double) * s->C;
    const bool contiguous = ld == s->C && s->C != s->shards[0].ld;
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        const bool local = s->multiprocess || s->shards.size() == 1;
        const double *rows = local ? tab : tab + sh.row0 * ld;       // this shard's constraint rows
        const double *obj = local ? tab + (int64_t)sh.m_local * ld : tab + s->m * ld;
        if (contiguous) {
            // one contiguous DMA per shard (a pitched 2-D copy of odd-width rows runs at a fraction
            // of PCIe speed), staged in the spare ping-pong buffer, then re-pitched on the device
            const int other = sh.cur ^ 1;
            if (!sh.tabs[other])
                CU_TRY(cudaMalloc(&sh.tabs[other], sizeof(double) * sh.cap_ld * sh.cap_rows));
            double *stage = sh.tabs[other];
            if (sh.m_local > 0)
                CU_TRY(cudaMemcpyAsync(stage, rows, wbytes * sh.m_local, cudaMemcpyHostToDevice, sh.stream));
            CU_TRY(cudaMemcpyAsync(stage + (int64_t)sh.m_local * s->C, obj, wbytes,
                                   cudaMemcpyHostToDevice, sh.stream));
            k_repitch<<<148 * 8, 256, 0, sh.stream>>>(stage, s->C, sh.tab, sh.ld, sh.R_local);
        } else {
            if (sh.m_local > 0)
                CU_TRY(cudaMemcpy2DAsync(sh.tab, sizeof(double) * sh.ld, rows, sizeof(double) * ld,
                                         wbytes, sh.m_local, cudaMemcpyHostToDevice, sh.stream));
            CU_TRY(cudaMemcpyAsync(sh.tab + (int64_t)sh.m_local * sh.ld, obj, wbytes,
                                   cudaMemcpyHostToDevice, sh.stream));
            if (sh.ld > s->C)
                k_zero_pad<<<64, 256, 0, sh.stream>>>(sh.tab, sh.ld, sh.R_local, (int)s->C);
        }
        if (basis && sh.m_local > 0)
            CU_TRY(cudaMemcpyAsync(sh.basis, basis + (local ? 0 : sh.row0), sizeof(int32_t) * sh.m_local,
                                   cudaMemcpyHostToDevice, sh.stream));
    }
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    CU_TRY(cudaGetLastError());
    s->iters_done = 0;
    return B200LP_OK;
}

static int download_locked(b200lp_solver *s, double *tab, int64_t ld, int32_t *basis)
{
    if (tab && ld < s->C) return fail(B200LP_ERR_INVALID_ARG, "download", "ld < C");
    const size_t wbytes = sizeof(double) * s->C;
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        const bool local = s->multiprocess || s->shards.size() == 1;
        if (tab) {
            if (local) {
                CU_TRY(cudaMemcpy2DAsync(tab, sizeof(double) * ld, sh.tab, sizeof(double) * sh.ld,
                                         wbytes, sh.R_local, cudaMemcpyDeviceToHost, sh.stream));
            } else {
                if (sh.m_local > 0)
                    CU_TRY(cudaMemcpy2DAsync(tab + sh.row0 * ld, sizeof(double) * ld, sh.tab,
                                             sizeof(double) * sh.ld, wbytes, sh.m_local,
                                             cudaMemcpyDeviceToHost, sh.stream));
                if (&sh == &s->shards[0])
                    CU_TRY(cudaMemcpyAsync(tab + s->m * ld, sh.tab + (int64_t)sh.m_local * sh.ld,
                                           wbytes, cudaMemcpyDeviceToHost, sh.stream));
            }
        }
        if (basis && sh.m_local > 0)
            CU_TRY(cudaMemcpyAsync(basis + (local ? 0 : sh.row0), sh.basis,
                                   sizeof(int32_t) * sh.m_local, cudaMemcpyDeviceToHost, sh.stream));
    }
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    return B200LP_OK;
}

static int download_solution_locked(b200lp_solver *s, double *rhs, double *obj_row, int32_t *basis)
{
    const bool local = s->multiprocess || s->shards.size() == 1;
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        if (rhs) {
            k_gather_col<<<(sh.R_local + 255) / 256, 256, 0, sh.stream>>>(sh.tab, sh.ld, sh.R_local,
                                                                          (int)(s->C - 1), sh.colout);
            s->kernel_launches++;
            if (local) {
                CU_TRY(cudaMemcpyAsync(rhs, sh.colout, sizeof(double) * sh.R_local,
                                       cudaMemcpyDeviceToHost, sh.stream));
            } else {
                if (sh.m_local > 0)
                    CU_TRY(cudaMemcpyAsync(rhs + sh.row0, sh.colout, sizeof(double) * sh.m_local,
                                           cudaMemcpyDeviceToHost, sh.stream));
                if (&sh == &s->shards[0])
                    CU_TRY(cudaMemcpyAsync(rhs + s->m, sh.colout + sh.m_local, sizeof(double),
                                           cudaMemcpyDeviceToHost, sh.stream));
            }
        }
        if (obj_row && &sh == &s->shards[0])
            CU_TRY(cudaMemcpyAsync(obj_row, sh.tab + (int64_t)sh.m_local * sh.ld,
                                   sizeof(double) * s->C, cudaMemcpyDeviceToHost, sh.stream));
        if (basis && sh.m_local > 0)
            CU_TRY(cudaMemcpyAsync(basis + (local ? 0 : sh.row0), sh.basis,
                                   sizeof(int32_t) * sh.m_local, cudaMemcpyDeviceToHost, sh.stream));
    }
    for (Shard &sh : s->shards) {
        CU_TRY(cudaSetDevice(sh.device));
        CU_TRY(cudaStreamSynchronize(sh.stream));
    }
    CU_TRY(cudaGetLastError());
    return B200LP_OK;
}

// Exchange mode for a sharded solver.  Preferred: peer-mapped buffers (mode 2) -- every rank maps
// every other rank's exchange buffer (same process: peer access; other processes: CUDA IPC
// handles, all-gathered over the NCCL communicator that exists anyway) and k_iter's look role
// moves the candidates itself.  Anything missing -> mode 1 (NCCL all-gather between two kernels).
// B200LP_EXCHANGE=nccl|p2p overrides (p2p: fail instead of falling back).
static int setup_exchange(b200lp_solver *s)
{
    s->xmode = 0;
    if (s->world == 1) return B200LP_OK;
    s->xmode = 1;
    const char *env = getenv("B200LP_EXCHANGE");
    const bool force_nccl = env && std::strcmp(env, "nccl") == 0;
    const bool force_p2p = env && std::strcmp(env, "p2p") == 0;
    if (const char *t = getenv("B200LP_PEER_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) s->peer_timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    if (force_nccl) return B200LP_OK;
    bool ok = true;
    std::string why;
    if (!s->multiprocess) {
        for (Shard &a : s->shards) {
            if (cudaSetDevice(a.device) != cudaSuccess) { ok = false; break; }
            for (Shard &b : s->shards) {
                if (&a == &b || a.device == b.device) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, a.device, b.device) != cudaSuccess || !can) {
                    ok = false; why = "no peer access between local devices"; break;
                }
                // enabled once per process and pair (asking again returns -- and a sanitizer
                // reports -- cudaErrorPeerAccessAlreadyEnabled)
                static bool enabled[kMaxDeviceLocks][kMaxDeviceLocks];
                static std::mutex enabled_mu;
                std::lock_guard<std::mutex> lk(enabled_mu);
                const bool tracked = a.device >= 0 && a.device < kMaxDeviceLocks && b.device >= 0 &&
                                     b.device < kMaxDeviceLocks;
                if (tracked && enabled[a.device][b.device]) continue;
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    ok = false; why = cudaGetErrorString(e); break;
                }
                (void)cudaGetLastError();
                if (tracked) enabled[a.device][b.device] = true;
            }
            if (!ok) break;
        }
        if (ok)
            for (Shard &a : s->shards)
                for (Shard &b : s->shards) a.xchg.peer[b.rank] = b.xbuf;
    } else {
        Shard &sh = s->shards[0];
        // all-gather (ok flag, IPC handle) through a small device buffer
        struct Blob { cudaIpcMemHandle_t h; int ok; int pad[15]; };
        static_assert(sizeof(Blob) == 128, "blob size");
        std::vector<Blob> all((size_t)s->world);
        Blob mine;
        std::memset(&mine, 0, sizeof(mine));
        mine.ok = cudaIpcGetMemHandle(&mine.h, sh.xbuf) == cudaSuccess ? 1 : 0;
        (void)cudaGetLastError();
        Blob *d = nullptr;
        CU_TRY(cudaMalloc(&d, sizeof(Blob) * (s->world + 1)));
        CU_TRY(cudaMemcpyAsync(d + s->world, &mine, sizeof(Blob), cudaMemcpyHostToDevice, sh.stream));
        ncclResult_t r = g_nccl.AllGather(d + s->world, d, sizeof(Blob), ncclChar, sh.comm, sh.stream);
        if (r != ncclSuccess) { cudaFree(d); return fail(B200LP_ERR_NCCL, "ncclAllGather(ipc)", g_nccl.GetErrorString(r)); }
        CU_TRY(cudaMemcpyAsync(all.data(), d, sizeof(Blob) * s->world, cudaMemcpyDeviceToHost, sh.stream));
        CU_TRY(cudaStreamSynchronize(sh.stream));
        cudaFree(d);
        for (int g = 0; g < s->world; ++g) ok = ok && all[(size_t)g].ok;
        int opened_ok = 1;
        if (ok) {
            for (int g = 0; g < s->world && opened_ok; ++g) {
                if (g == sh.rank) { sh.xchg.peer[g] = sh.xbuf; continue; }
                void *ptr = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[(size_t)g].h, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    (void)cudaGetLastError();
                    why = cudaGetErrorString(e);
                    opened_ok = 0;
                } else {
                    sh.xchg.peer[g] = static_cast<double *>(ptr);
                    sh.ipc_opened[g] = true;
                }
            }
        } else {
            why = "cudaIpcGetMemHandle failed on some rank";
            opened_ok = 0;
        }
        // every rank must take the same path: agree (min) on success through one more all-gather
        int *dflag = nullptr;
        CU_TRY(cudaMalloc(&dflag, sizeof(int) * (s->world + 1)));
        CU_TRY(cudaMemcpyAsync(dflag + s->world, &opened_ok, sizeof(int), cudaMemcpyHostToDevice, sh.stream));
        r = g_nccl.AllGather(dflag + s->world, dflag, sizeof(int), ncclChar, sh.comm, sh.stream);
        if (r != ncclSuccess) { cudaFree(dflag); return fail(B200LP_ERR_NCCL, "ncclAllGather(ipc ok)", g_nccl.GetErrorString(r)); }
        std::vector<int> flags((size_t)s->world);
        CU_TRY(cudaMemcpyAsync(flags.data(), dflag, sizeof(int) * s->world, cudaMemcpyDeviceToHost, sh.stream));
        CU_TRY(cudaStreamSynchronize(sh.stream));
        cudaFree(dflag);
        ok = true;
        for (int g = 0; g < s->world; ++g) ok = ok && flags[(size_t)g];
    }
    if (ok) {
        s->xmode = 2;
        return B200LP_OK;
    }
    if (force_p2p) return fail(B200LP_ERR_CUDA, "peer-mapped exchange unavailable", why.c_str());
    return B200LP_OK;
}

static int create_common(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max,
                         b200lp_solver **out)
{
    if (!out) return fail(B200LP_ERR_INVALID_ARG, "create", "null out pointer");
    *out = nullptr;
    if (R < 2 || C < 2 || R > (1ll << 30) || C > (1ll << 30))
        return fail(B200LP_ERR_INVALID_ARG, "create", "need R >= 2, C >= 2");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        (void)cudaGetLastError();
        return fail(B200LP_ERR_NO_DEVICE, "cudaGetDeviceCount",
                    e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device");
    }
    b200lp_solver *s = new (std::nothrow) b200lp_solver();
    if (!s) return fail(B200LP_ERR_OUT_OF_MEMORY, "create", "host allocation");
    if (opts) s->opts = *opts; else std::memset(&s->opts, 0, sizeof(s->opts));
    s->R = R; s->C = C; s->m = R - 1; s->is_max = is_max ? 1 : 0;
    fill_thresholds(s);
    if (const char *t = getenv("B200LP_SPIN_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) s->spin_timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    *out = s;
    return B200LP_OK;
}

static int ensure_nccl()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    const char *why = "";
    if (!g_nccl.load(&why)) return fail(B200LP_ERR_NCCL, "NCCL", why);
    return B200LP_OK;
}

// ---- handle pool of the one-shot calls ----------------------------------------------------------
static std::mutex g_pool_mu;
static std::vector<b200lp_solver *> g_pool;            // idle single-GPU handles
static constexpr size_t kPoolSlots = 4;

// Re-bind an idle pooled handle to a new problem shape and option set.
static int rebind(b200lp_solver *s, const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max)
{
    if (opts) s->opts = *opts; else std::memset(&s->opts, 0, sizeof(s->opts));
    s->R = R; s->C = C; s->m = R - 1; s->is_max = is_max ? 1 : 0;
    fill_thresholds(s);
    s->iters_done = 0;
    Shard &sh = s->shards[0];
    sh.row0 = 0; sh.m_local = (int)(R - 1); sh.R_local = (int)R;
    sh.ld = round_up(C, 16);
    shape_shard(sh);
    CU_TRY(cudaSetDevice(sh.device));
    return ensure_trace(s, sh);
}

// A handle for a one-shot solve on one GPU: from the pool when an idle one is large enough (and
// not wastefully larger), else freshly created.
static int acquire(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max, b200lp_solver **out)
{
    const int nd = opts ? std::max(1, (int)opts->ndev) : 1;
    const int device = opts ? opts->devices[0] : 0;
    if (nd == 1 && R >= 2 && C >= 2) {
        const int64_t ld = round_up(C, 16);
        std::unique_lock<std::mutex> lk(g_pool_mu);
        for (size_t k = 0; k < g_pool.size(); ++k) {
            b200lp_solver *c = g_pool[k];
            const Shard &sh = c->shards[0];
            if (sh.device != device || R > sh.cap_rows || ld > sh.cap_ld) continue;
            if (sh.cap_rows * sh.cap_ld > 4 * std::max<int64_t>(R * ld, 64 * 256)) continue;
            g_pool.erase(g_pool.begin() + (long)k);
            lk.unlock();
            const int rc = rebind(c, opts, R, C, is_max);
            if (rc != B200LP_OK) { b200lp_destroy(c); return rc; }
            *out = c;
            return B200LP_OK;
        }
    }
    return b200lp_create(opts, R, C, is_max, out);
}

static void release(b200lp_solver *s)
{
    if (!s) return;
    if (s->shards.size() == 1 && !s->multiprocess && poolable(s, s->shards[0])) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < kPoolSlots) { g_pool.push_back(s); return; }
    }
    b200lp_destroy(s);
}

// ---- single-CTA path for tableaus that fit in shared memory (small.cuh) ----------------------------
// One context per device: a stream and a mapped pinned buffer the kernel reads its input from and
// writes its results to (zero-copy: one launch + one synchronise per solve, no cudaMemcpy calls).
struct SmallCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    unsigned char *h_buf = nullptr;
    size_t cap = 0;
    bool attr_set = false;
};
static SmallCtx g_small[kMaxDeviceLocks];
static constexpr size_t kSmallSmemMax = 225 * 1024;

static bool small_enabled()
{
    const char *e = getenv("B200LP_SMALL");          // B200LP_SMALL=0: always take the big path
    return !(e && e[0] == '0');
}

static bool small_ok(const b200lp_opts *o, int64_t R, int64_t C_in, int64_t C_main)
{
    if (!small_enabled()) return false;
    if (o && (o->ndev > 1 || o->pivot_variant != 0 || o->poll_interval != 0 || o->time_kernels != 0))
        return false;
    if (R < 2 || C_main < 2 || C_in < C_main || R > 8192 || C_in > 28000) return false;
    return small_smem_bytes((int)R, (int)C_in, (int)C_main) <= kSmallSmemMax;
}

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// b200lp_solve (art_tab == nullptr) or b200lp_solve_two_phase on one CTA.
static int solve_small(const b200lp_opts *opts, double *art_tab, int64_t C_art, int64_t ld_art,
                       int32_t *art_basis, double *tab, int64_t R, int64_t C, int64_t ld,
                       int32_t *basis, int32_t is_max, b200lp_result *out, int32_t *trace_j,
                       int32_t *trace_r)
{
    const double t0 = now_ms();
    const bool two = art_tab != nullptr;
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev <= 0) {
        (void)cudaGetLastError();
        return fail(B200LP_ERR_NO_DEVICE, "cudaGetDeviceCount",
                    e0 != cudaSuccess ? cudaGetErrorString(e0) : "no CUDA device");
    }
    const int device = opts ? opts->devices[0] : 0;
    SmallCtx &cx = g_small[((device % kMaxDeviceLocks) + kMaxDeviceLocks) % kMaxDeviceLocks];
    std::lock_guard<std::mutex> lk(cx.mu);
    CU_TRY(cudaSetDevice(device));
    const bool wb = opts && opts->writeback_full;
    const int tcap = (!two && opts && opts->trace_capacity > 0) ? opts->trace_capacity : 0;
    const int64_t C_in = two ? C_art : C;
    // buffer layout
    size_t off = 0;
    const size_t o_in = off;      off = align16(off + sizeof(double) * R * C_in);
    const size_t o_inb = off;     off = align16(off + sizeof(int32_t) * (R - 1));
    const size_t o_mobj = off;    off = align16(off + sizeof(double) * C);
    const size_t o_hdr = off;     off = align16(off + sizeof(SmallHeader));
    const size_t o_rhs = off;     off = align16(off + sizeof(double) * R);
    const size_t o_obj = off;     off = align16(off + sizeof(double) * C);
    const size_t o_basis = off;   off = align16(off + sizeof(int32_t) * (R - 1));
    const size_t o_full = off;    off = align16(off + (wb ? sizeof(double) * R * C : 0));
    const size_t o_afull = off;   off = align16(off + (wb && two ? sizeof(double) * R * C_in : 0));
    const size_t o_abasis = off;  off = align16(off + (wb && two ? sizeof(int32_t) * (R - 1) : 0));
    const size_t o_trace = off;   off = align16(off + sizeof(int2) * (size_t)tcap);
    if (off > cx.cap) {
        if (cx.h_buf) cudaFreeHost(cx.h_buf);
        cx.h_buf = nullptr; cx.cap = 0;
        const size_t want = std::max<size_t>(off, 1u << 20);
        CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&cx.h_buf), want, cudaHostAllocMapped));
        cx.cap = want;
    }
    if (!cx.stream) CU_TRY(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
    if (!cx.attr_set) {
        CU_TRY(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmallSmemMax));
        cx.attr_set = true;
    }
    unsigned char *hb = cx.h_buf;
    unsigned char *db = nullptr;
    CU_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&db), hb, 0));
    // stage the input (packed rows)
    {
        const double *src = two ? art_tab : tab;
        const int64_t lds = two ? ld_art : ld;
        double *dst = reinterpret_cast<double *>(hb + o_in);
        for (int64_t r = 0; r < R; ++r) std::memcpy(dst + r * C_in, src + r * lds, sizeof(double) * C_in);
        std::memcpy(hb + o_inb, two ? art_basis : basis, sizeof(int32_t) * (R - 1));
        if (two) std::memcpy(hb + o_mobj, tab + (R - 1) * ld, sizeof(double) * C);
        reinterpret_cast<SmallHeader *>(hb + o_hdr)->status = ST_RUNNING;
    }
    double tol = (opts && opts->fp_tolerance_factor > 0) ? opts->fp_tolerance_factor : 1024.0;
    SmallArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in_tab = reinterpret_cast<const double *>(db + o_in);
    a.in_basis = reinterpret_cast<const int32_t *>(db + o_inb);
    a.main_obj = reinterpret_cast<const double *>(db + o_mobj);
    a.R = (int)R; a.C_in = (int)C_in; a.C_main = (int)C;
    a.two_phase = two ? 1 : 0; a.is_max = is_max ? 1 : 0; a.rule = opts ? opts->pivot_rule : 0;
    a.feas_reference = (opts && opts->feas_mode == B200LP_FEAS_REFERENCE) ? 1 : 0;
    b200lp_thresholds(tol, &a.thr_enter, &a.thr_pivot, &a.thr_feas);
    a.max_iters = opts ? opts->max_iters : 0;
    a.hdr = reinterpret_cast<SmallHeader *>(db + o_hdr);
    a.out_rhs = reinterpret_cast<double *>(db + o_rhs);
    a.out_obj = reinterpret_cast<double *>(db + o_obj);
    a.out_basis = reinterpret_cast<int32_t *>(db + o_basis);
    a.out_full = wb ? reinterpret_cast<double *>(db + o_full) : nullptr;
    a.out_art_full = (wb && two) ? reinterpret_cast<double *>(db + o_afull) : nullptr;
    a.out_art_basis = (wb && two) ? reinterpret_cast<int32_t *>(db + o_abasis) : nullptr;
    a.trace = tcap ? reinterpret_cast<int2 *>(db + o_trace) : nullptr;
    a.trace_cap = tcap;
    const size_t smem = small_smem_bytes((int)R, (int)C_in, (int)C);
    const double t1 = now_ms();
    k_small<<<1, kSmallThreads, smem, cx.stream>>>(a);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(cx.stream));
    const double t2 = now_ms();
    const SmallHeader h = *reinterpret_cast<const SmallHeader *>(hb + o_hdr);
    if (h.status == ST_RUNNING) return fail(B200LP_ERR_INTERNAL, "solve", "k_small left no verdict");
    const int status = h.status;
    int tl = 0;
    const bool have = h.wrote != 0;    // the solve reached (and solved) the tableau the caller reads back
    if (have) {
        const double *rhs = reinterpret_cast<const double *>(hb + o_rhs);
        const double *obj = reinterpret_cast<const double *>(hb + o_obj);
        if (wb) {
            const double *full = reinterpret_cast<const double *>(hb + o_full);
            for (int64_t r = 0; r < R; ++r) std::memcpy(tab + r * ld, full + r * C, sizeof(double) * C);
            if (two) {
                const double *af = reinterpret_cast<const double *>(hb + o_afull);
                for (int64_t r = 0; r < R; ++r)
                    std::memcpy(art_tab + r * ld_art, af + r * C_in, sizeof(double) * C_in);
                std::memcpy(art_basis, hb + o_abasis, sizeof(int32_t) * (R - 1));
            }
        } else {
            for (int64_t i = 0; i < R; ++i) tab[i * ld + (C - 1)] = rhs[i];
            std::memcpy(tab + (R - 1) * ld, obj, sizeof(double) * C);
        }
        std::memcpy(basis, hb + o_basis, sizeof(int32_t) * (R - 1));
        if (tcap) {
            tl = (int)std::min<long long>(h.iters, tcap);
            const int2 *tr = reinterpret_cast<const int2 *>(hb + o_trace);
            for (int q = 0; q < tl; ++q) {
                if (trace_j) trace_j[q] = tr[q].x;
                if (trace_r) trace_r[q] = tr[q].y;
            }
        }
    }
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->status = status;
        out->n_devices = 1;
        out->iterations = h.iters;
        out->iterations_phase1 = h.iters_phase1;
        out->iterations_cleanup = h.iters_cleanup;
        out->redundant_rows = h.redundant;
        out->objective = h.objective;
        out->ms_h2d = t1 - t0;
        out->ms_solve = t2 - t1;
        out->kernel_launches = 1;
        out->loop_mode = 3;
        out->h2d_bytes = (int64_t)(sizeof(double) * R * C_in + sizeof(int32_t) * (R - 1) + (two ? sizeof(double) * C : 0));
        out->d2h_bytes = (int64_t)(sizeof(SmallHeader) + sizeof(double) * (R + C) + sizeof(int32_t) * (R - 1) +
                                   (wb ? sizeof(double) * R * C : 0) + sizeof(int2) * (size_t)tl);
        out->bytes_per_pivot = 16 * R * C;
        out->trace_len = tl;
        out->ms_total = now_ms() - t0;
    }
    return status;
}

} // namespace b200lp

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

void b200lp_thresholds(double tol, double *enter, double *pivot, double *feas)
{
    if (!(tol > 0)) tol = 1024.0;
    if (enter) *enter = (tol / 8.0) * kClEps;   // src/simplex.lisp:370-371, 377-378
    if (pivot) *pivot = (tol / 2.0) * kClEps;   // src/simplex.lisp:386-387
    if (feas) *feas = tol * kClEps;             // src/simplex.lisp:405-406
}

void b200lp_partition(int64_t m, int32_t nranks, int32_t rank, int64_t *row_begin, int64_t *row_end)
{
    if (nranks < 1) nranks = 1;
    const int64_t per = (m + nranks - 1) / nranks;
    int64_t b = std::min<int64_t>(m, per * rank);
    int64_t e = std::min<int64_t>(m, b + per);
    if (row_begin) *row_begin = b;
    if (row_end) *row_end = e;
}

int b200lp_version(void) { return B200LP_VERSION; }

void b200lp_abi_sizes(int64_t *opts_size, int64_t *result_size)
{
    if (opts_size) *opts_size = (int64_t)sizeof(b200lp_opts);
    if (result_size) *result_size = (int64_t)sizeof(b200lp_result);
}

void b200lp_shutdown(void)
{
    for (int d = 0; d < kMaxDeviceLocks; ++d) {
        SmallCtx &cx = g_small[d];
        std::lock_guard<std::mutex> lk(cx.mu);
        if (cx.h_buf || cx.stream) {
            cudaSetDevice(d);
            if (cx.stream) { cudaStreamSynchronize(cx.stream); cudaStreamDestroy(cx.stream); }
            if (cx.h_buf) cudaFreeHost(cx.h_buf);
            cx.stream = nullptr; cx.h_buf = nullptr; cx.cap = 0;
        }
    }
    std::vector<b200lp_solver *> idle;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        idle.swap(g_pool);
    }
    for (b200lp_solver *s : idle) b200lp_destroy(s);
}

int b200lp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

const char *b200lp_last_error(void) { return g_last_error.c_str(); }

const char *b200lp_strerror(int code)
{
    switch (code) {
    case B200LP_OK: return "optimal";
    case B200LP_UNBOUNDED: return "Problem is unbounded";
    case B200LP_INFEASIBLE: return "Problem has no feasible region";
    case B200LP_ITERATION_LIMIT: return "iteration limit reached";
    case B200LP_ARTIFICIAL_STUCK: return "Artificial variable still in basis and cannot be replaced";
    case B200LP_ARTIFICIAL_NONZERO: return "Artificial variable still non-zero";
    case B200LP_ERR_INVALID_ARG: return "invalid argument";
    case B200LP_ERR_CUDA: return "CUDA error";
    case B200LP_ERR_NCCL: return "NCCL error";
    case B200LP_ERR_NO_DEVICE: return "no CUDA device";
    case B200LP_ERR_OUT_OF_MEMORY: return "out of device memory";
    case B200LP_ERR_INTERNAL: return "internal error";
    case B200LP_ERR_PEER_TIMEOUT: return "timed out waiting for a peer GPU";
    default: return "unknown status";
    }
}

int b200lp_create(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max,
                  b200lp_solver **out)
{
    try {
        RC_TRY(create_common(opts, R, C, is_max, out));
        b200lp_solver *s = *out;
        const int nd = std::max(1, (int)s->opts.ndev);
        if (nd > B200LP_MAX_DEVICES || nd > s->m) {
            b200lp_destroy(s); *out = nullptr;
            return fail(B200LP_ERR_INVALID_ARG, "create", "ndev out of range");
        }
        s->world = nd;
        s->multiprocess = false;
        s->shards.resize(nd);
        for (int g = 0; g < nd; ++g) {
            Shard &sh = s->shards[g];
            sh.device = s->opts.devices[g];
            if (s->opts.ndev == 0) sh.device = s->opts.devices[0];
            sh.rank = g;
            int64_t b, e;
            b200lp_partition(s->m, nd, g, &b, &e);
            sh.row0 = b; sh.m_local = (int)(e - b); sh.R_local = sh.m_local + 1;
        }
        int rc = B200LP_OK;
        for (Shard &sh : s->shards) { rc = alloc_shard(s, sh); if (rc) break; }
        // Several GPUs in one process: peer access is all the in-kernel exchange needs; an NCCL
        // communicator (seconds to create) is only built when peers cannot be mapped.
        if (!rc) rc = setup_exchange(s);
        if (!rc && nd > 1 && s->xmode == 1) {
            rc = ensure_nccl();
            if (!rc) {
                std::vector<ncclComm_t> comms(nd);
                std::vector<int> devs(nd);
                for (int g = 0; g < nd; ++g) devs[g] = s->shards[g].device;
                ncclResult_t r = g_nccl.CommInitAll(comms.data(), nd, devs.data());
                if (r != ncclSuccess) rc = fail(B200LP_ERR_NCCL, "ncclCommInitAll", g_nccl.GetErrorString(r));
                else for (int g = 0; g < nd; ++g) s->shards[g].comm = comms[g];
            }
        }
        if (rc) { b200lp_destroy(s); *out = nullptr; return rc; }
        return B200LP_OK;
    } catch (const std::exception &ex) {
        return fail(B200LP_ERR_INTERNAL, "create", ex.what());
    }
}

int b200lp_comm_unique_id(void *unique_id_128)
{
    if (!unique_id_128) return fail(B200LP_ERR_INVALID_ARG, "comm_unique_id", "null");
    RC_TRY(ensure_nccl());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(unique_id_128, &id, 128);
    return B200LP_OK;
}

int b200lp_create_sharded(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max,
                          int32_t rank, int32_t nranks, const void *unique_id_128,
                          b200lp_solver **out)
{
    try {
        if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !unique_id_128))
            return fail(B200LP_ERR_INVALID_ARG, "create_sharded", "bad rank/nranks/id");
        RC_TRY(create_common(opts, R, C, is_max, out));
        b200lp_solver *s = *out;
        if (nranks > s->m) {
            b200lp_destroy(s); *out = nullptr;
            return fail(B200LP_ERR_INVALID_ARG, "create_sharded", "more ranks than rows");
        }
        s->world = nranks;
        s->multiprocess = true;
        s->shards.resize(1);
        Shard &sh = s->shards[0];
        sh.device = s->opts.devices[0];
        sh.rank = rank;
        int64_t b, e;
        b200lp_partition(s->m, nranks, rank, &b, &e);
        sh.row0 = b; sh.m_local = (int)(e - b); sh.R_local = sh.m_local + 1;
        int rc = alloc_shard(s, sh);
        if (!rc && nranks > 1) {
            rc = ensure_nccl();
            if (!rc) {
                ncclUniqueId id;
                std::memcpy(&id, unique_id_128, 128);
                ncclResult_t r = g_nccl.CommInitRank(&sh.comm, nranks, id, rank);
                if (r != ncclSuccess) rc = fail(B200LP_ERR_NCCL, "ncclCommInitRank", g_nccl.GetErrorString(r));
            }
        }
        if (!rc) rc = setup_exchange(s);
        if (rc) { b200lp_destroy(s); *out = nullptr; return rc; }
        return B200LP_OK;
    } catch (const std::exception &ex) {
        return fail(B200LP_ERR_INTERNAL, "create_sharded", ex.what());
    }
}

int b200lp_shard_rows(const b200lp_solver *s, int64_t *row_begin, int64_t *row_end)
{
    if (!s || s->shards.empty()) return fail(B200LP_ERR_INVALID_ARG, "shard_rows", "null solver");
    if (row_begin) *row_begin = s->shards.front().row0;
    if (row_end) *row_end = s->shards.back().row0 + s->shards.back().m_local;
    return B200LP_OK;
}

void b200lp_destroy(b200lp_solver *s)
{
    if (!s) return;
    for (Shard &sh : s->shards) free_shard(sh);
    delete s;
}

int b200lp_upload(b200lp_solver *s, const double *tab, int64_t ld, const int32_t *basis)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "upload", "null solver");
    std::lock_guard<std::mutex> lk(s->mu);
    return upload_locked(s, tab, ld, basis);
}

int b200lp_download(b200lp_solver *s, double *tab, int64_t ld, int32_t *basis)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "download", "null solver");
    std::lock_guard<std::mutex> lk(s->mu);
    return download_locked(s, tab, ld, basis);
}

int b200lp_download_solution(b200lp_solver *s, double *rhs, double *obj_row, int32_t *basis)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "download_solution", "null solver");
    std::lock_guard<std::mutex> lk(s->mu);
    return download_solution_locked(s, rhs, obj_row, basis);
}

int b200lp_find_entering_column(b200lp_solver *s, int64_t *col)
{
    if (!s || !col) return fail(B200LP_ERR_INVALID_ARG, "find_entering_column", "null");
    std::lock_guard<std::mutex> lk(s->mu);
    RC_TRY(push_state(s, ST_RUNNING, 0, -1, -1));
    for (Shard &sh : s->shards) { CU_TRY(cudaSetDevice(sh.device)); launch_enter(s, sh); }
    DevState st;
    RC_TRY(pull_state(s, &st));
    *col = (st.status == ST_RUNNING) ? st.j : -1;
    return B200LP_OK;
}

int b200lp_find_pivoting_row(b200lp_solver *s, int64_t j, int64_t *row)
{
    if (!s || !row) return fail(B200LP_ERR_INVALID_ARG, "find_pivoting_row", "null");
    if (j < 0 || j >= s->C) return fail(B200LP_ERR_INVALID_ARG, "find_pivoting_row", "column out of range");
    std::lock_guard<std::mutex> lk(s->mu);
    RC_TRY(push_state(s, ST_RUNNING, 0, (int)j, -1));
    for (Shard &sh : s->shards) { CU_TRY(cudaSetDevice(sh.device)); launch_ratio(s, sh, false); }
    if (s->world > 1) {
        // candidates travel with their header only being meaningful; rows may be stale
        RC_TRY(exchange(s));
        for (Shard &sh : s->shards) { CU_TRY(cudaSetDevice(sh.device)); launch_winner(s, sh, false); }
    }
    DevState st;
    RC_TRY(pull_state(s, &st));
    *row = (st.status == ST_RUNNING) ? st.p : -1;
    return B200LP_OK;
}

int b200lp_pivot(b200lp_solver *s, int64_t j, int64_t r)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "pivot", "null solver");
    if (j < 0 || j >= s->C || r < 0 || r >= s->m)
        return fail(B200LP_ERR_INVALID_ARG, "pivot", "row/column out of range");
    std::lock_guard<std::mutex> lk(s->mu);
    RC_TRY(push_state(s, ST_RUNNING, 0, (int)j, (int)r));
    RC_TRY(enqueue_iteration(s, false, false, false));
    DevState st;
    RC_TRY(pull_state(s, &st));
    s->iters_done = st.iters;
    return B200LP_OK;
}

int b200lp_iterate(b200lp_solver *s, int64_t max_iters, b200lp_result *out, int32_t *trace_j,
                   int32_t *trace_r)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "iterate", "null solver");
    std::lock_guard<std::mutex> lk(s->mu);
    std::vector<int> devs;
    for (const Shard &sh : s->shards) devs.push_back(((sh.device % kMaxDeviceLocks) + kMaxDeviceLocks) % kMaxDeviceLocks);
    std::sort(devs.begin(), devs.end());
    devs.erase(std::unique(devs.begin(), devs.end()), devs.end());
    for (int d : devs) g_device_mu[d].lock();              // ascending order: no lock cycles
    struct Unlock {
        std::vector<int> &d;
        ~Unlock() { for (auto it = d.rbegin(); it != d.rend(); ++it) g_device_mu[*it].unlock(); }
    } unlock{devs};
    try {
        return iterate_locked(s, max_iters, out, trace_j, trace_r);
    } catch (const std::exception &ex) {
        return fail(B200LP_ERR_INTERNAL, "iterate", ex.what());
    }
}

int b200lp_set_time_kernels(b200lp_solver *s, int32_t on)
{
    if (!s) return fail(B200LP_ERR_INVALID_ARG, "set_time_kernels", "null solver");
    std::lock_guard<std::mutex> lk(s->mu);
    s->opts.time_kernels = on ? 1 : 0;
    return B200LP_OK;
}

int b200lp_solve(const b200lp_opts *opts, double *tab, int64_t R, int64_t C, int64_t ld,
                 int32_t *basis, int32_t is_max, b200lp_result *out, int32_t *trace_j,
                 int32_t *trace_r)
{
    if (!tab || !basis || ld < C) return fail(B200LP_ERR_INVALID_ARG, "solve", "null buffer or ld < C");
    if (small_ok(opts, R, C, C)) {
        try {
            return solve_small(opts, nullptr, 0, 0, nullptr, tab, R, C, ld, basis, is_max, out, trace_j, trace_r);
        } catch (const std::exception &ex) {
            return fail(B200LP_ERR_INTERNAL, "solve", ex.what());
        }
    }
    const double t0 = now_ms();
    b200lp_solver *s = nullptr;
    RC_TRY(acquire(opts, R, C, is_max, &s));
    b200lp_result res;
    std::memset(&res, 0, sizeof(res));
    int rc = B200LP_OK;
    double t_h2d = now_ms();
    rc = b200lp_upload(s, tab, ld, basis);
    const double ms_h2d = now_ms() - t_h2d;
    int status = rc;
    if (rc == B200LP_OK) {
        status = b200lp_iterate(s, 0, &res, trace_j, trace_r);
        if (status >= 0) {
            const double t_d2h = now_ms();
            if (s->opts.writeback_full) {
                rc = b200lp_download(s, tab, ld, basis);
                res.d2h_bytes += (int64_t)sizeof(double) * R * C + (int64_t)sizeof(int32_t) * (R - 1);
            } else {
                // only what the accessors read (src/simplex.lisp:74-120): RHS column, objective
                // row, basis.  Staged through contiguous host buffers, then scattered into `tab`.
                std::vector<double> rhs((size_t)R), obj((size_t)C);
                rc = b200lp_download_solution(s, rhs.data(), obj.data(), basis);
                if (rc == B200LP_OK) {
                    for (int64_t i = 0; i < R; ++i) tab[i * ld + (C - 1)] = rhs[(size_t)i];
                    std::memcpy(tab + (R - 1) * ld, obj.data(), sizeof(double) * C);
                }
                res.d2h_bytes += (int64_t)sizeof(double) * (R + C) + (int64_t)sizeof(int32_t) * (R - 1);
            }
            res.ms_d2h = now_ms() - t_d2h;
            if (rc != B200LP_OK) status = rc;
        }
    }
    if (status < 0) b200lp_destroy(s);                     // never pool a handle that saw an error
    else release(s);
    res.status = status;
    res.ms_h2d = ms_h2d;
    res.h2d_bytes = (int64_t)sizeof(double) * R * C + (int64_t)sizeof(int32_t) * (R - 1);
    res.ms_total = now_ms() - t0;
    if (out) *out = res;
    return status;
}

// Two-phase driver: src/simplex.lisp:402-452.  Single device (the transition's objective
// re-pricing is order dependent across all rows; BASELINE's sharded configs never need it).
int b200lp_solve_two_phase(const b200lp_opts *opts, double *art_tab, int64_t C_art, int64_t ld_art,
                           int32_t *art_basis, double *main_tab, int64_t R, int64_t C, int64_t ld,
                           int32_t *main_basis, int32_t is_max, b200lp_result *out)
{
    if (!art_tab || !art_basis || !main_tab || !main_basis || ld_art < C_art || ld < C || C_art < C)
        return fail(B200LP_ERR_INVALID_ARG, "solve_two_phase", "null buffer or bad dimensions");
    if (small_ok(opts, R, C_art, C)) {
        try {
            return solve_small(opts, art_tab, C_art, ld_art, art_basis, main_tab, R, C, ld, main_basis,
                               is_max, out, nullptr, nullptr);
        } catch (const std::exception &ex) {
            return fail(B200LP_ERR_INTERNAL, "solve_two_phase", ex.what());
        }
    }
    // Sharded (opts->ndev > 1, row blocks over the GPUs of this process): phase 1 and phase 2 are
    // ordinary sharded solves; the transition between them is a handful of small steps:
    //   * clean-up pivots (:419-434) are chosen on the shard that owns the row and applied with the
    //     sharded n-pivot-row (b200lp_pivot);
    //   * the coefficient copy (:437-441) is local to every shard;
    //   * the objective re-pricing (:444-451) subtracts scale_i * row_i for i = 0..m-1 IN ROW ORDER
    //     (two roundings per term): the scales come from the original objective row every shard
    //     holds a replica of, then the row is handed from shard to shard in rank order -- each
    //     applies its rows in order -- and the result is copied back to every replica.
    const double t0 = now_ms();
    b200lp_solver *art = nullptr, *mn = nullptr;
    b200lp_result r1, r2;
    std::memset(&r1, 0, sizeof(r1));
    std::memset(&r2, 0, sizeof(r2));
    int64_t cleanup = 0, redundant = 0;
    int status = B200LP_OK;
    std::vector<unsigned char *> d_is_basic;
    std::vector<int *> d_newcol;
    std::vector<double *> d_scales;
    auto finish = [&](int st) {
        if (art) {
            for (size_t g = 0; g < art->shards.size(); ++g) {
                cudaSetDevice(art->shards[g].device);
                if (g < d_is_basic.size()) cudaFree(d_is_basic[g]);
                if (g < d_newcol.size()) cudaFree(d_newcol[g]);
                if (g < d_scales.size()) cudaFree(d_scales[g]);
            }
            cudaSetDevice(art->shards[0].device);
        }
        if (st < 0) { b200lp_destroy(art); b200lp_destroy(mn); }
        else { release(art); release(mn); }
        if (out) {
            *out = r2;
            out->status = st;
            if (art) out->n_devices = art->world;
            out->iterations_phase1 = r1.iterations;
            out->iterations_cleanup = cleanup;
            out->redundant_rows = redundant;
            out->ms_solve = r1.ms_solve + r2.ms_solve;
            out->kernel_launches = r1.kernel_launches + r2.kernel_launches;
            out->h2d_bytes = (int64_t)sizeof(double) * R * (C + C_art);
            out->ms_total = now_ms() - t0;
        }
        return st;
    };
    const int64_t m = R - 1, nv = C - 1, art_nv = C_art - 1;
    // "zero" in the transition: the reference's absolute tests, or (default) scaled by the size of
    // the numbers that cancel in the phase-1 objective -- see B200LP_FEAS_* in the header
    const bool ref_feas = opts && opts->feas_mode == B200LP_FEAS_REFERENCE;
    const double obj0 = art_tab[(R - 1) * ld_art + art_nv];
    const double feas_scale = ref_feas ? 1.0 : std::fmax(1.0, std::fabs(obj0));
    int rc = acquire(opts, R, C_art, /*is_max=*/0, &art);         // phase 1 is a `min` problem
    if (rc) return finish(rc);
    rc = acquire(opts, R, C, is_max, &mn);
    if (rc) return finish(rc);
    if ((rc = b200lp_upload(art, art_tab, ld_art, art_basis))) return finish(rc);
    if ((rc = b200lp_upload(mn, main_tab, ld, main_basis))) return finish(rc);

    status = b200lp_iterate(art, 0, &r1, nullptr, nullptr);
    if (status != B200LP_OK) return finish(status);
    // (unless (fp= 0 obj tol) (error 'infeasible-problem-error))  :405-407
    const double thr_feas = art->thr_feas * feas_scale;
    if (!(std::fabs(0.0 - r1.objective) <= thr_feas)) return finish(B200LP_INFEASIBLE);

    const size_t nsh = art->shards.size();                // both handles share the partition
    d_is_basic.assign(nsh, nullptr);
    d_newcol.assign(nsh, nullptr);
    d_scales.assign(nsh, nullptr);
    std::vector<int32_t> hb((size_t)m);
#define TP_CU(expr)                                                                     \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            fail(B200LP_ERR_CUDA, #expr, cudaGetErrorString(e__));                      \
            return finish(B200LP_ERR_CUDA);                                             \
        }                                                                               \
    } while (0)
    for (Shard &as : art->shards) {
        TP_CU(cudaSetDevice(as.device));
        if (as.m_local > 0)
            TP_CU(cudaMemcpy(hb.data() + as.row0, as.basis, sizeof(int32_t) * as.m_local, cudaMemcpyDeviceToHost));
    }
    // drive zero-level artificial variables out of the basis  :419-434
    bool any_art = false;
    for (int64_t i = 0; i < m; ++i) any_art |= hb[(size_t)i] >= nv;
    if (any_art) {
        std::vector<unsigned char> is_basic((size_t)C_art);
        for (int64_t i = 0; i < m; ++i) {
            if (hb[(size_t)i] < nv) continue;
            size_t g = 0;
            while (g + 1 < nsh && i >= art->shards[g].row0 + art->shards[g].m_local) ++g;
            Shard &as = art->shards[g];                    // the shard that owns row i
            const int64_t li = i - as.row0;
            TP_CU(cudaSetDevice(as.device));
            if (!d_is_basic[g]) {
                TP_CU(cudaMalloc(&d_is_basic[g], (size_t)C_art));
                TP_CU(cudaMalloc(&d_newcol[g], sizeof(int)));
            }
            double rhs = 0.0;
            TP_CU(cudaMemcpy(&rhs, as.tab + li * as.ld + art_nv, sizeof(double), cudaMemcpyDeviceToHost));
            if (ref_feas ? (rhs != 0.0) : !(std::fabs(rhs) <= thr_feas))
                return finish(B200LP_ARTIFICIAL_NONZERO);
            std::fill(is_basic.begin(), is_basic.end(), 0);
            for (int64_t k = 0; k < m; ++k) is_basic[(size_t)hb[(size_t)k]] = 1;
            TP_CU(cudaMemcpy(d_is_basic[g], is_basic.data(), (size_t)C_art, cudaMemcpyHostToDevice));
            if (ref_feas)
                k_first_nonzero_nonbasic<<<1, 1024, 0, as.stream>>>(as.tab + li * as.ld, (int)nv,
                                                                    d_is_basic[g], d_newcol[g]);
            else
                k_largest_nonbasic<<<1, 1024, 0, as.stream>>>(as.tab + li * as.ld, (int)nv, d_is_basic[g],
                                                              art->thr_pivot, d_newcol[g]);
            int new_col = -1;
            TP_CU(cudaMemcpyAsync(&new_col, d_newcol[g], sizeof(int), cudaMemcpyDeviceToHost, as.stream));
            TP_CU(cudaStreamSynchronize(as.stream));
            if (new_col < 0) {
                if (ref_feas) return finish(B200LP_ARTIFICIAL_STUCK);
                ++redundant;                               // a combination of the other rows: keep it
                continue;
            }
            if ((rc = b200lp_pivot(art, new_col, i))) return finish(rc);
            hb[(size_t)i] = new_col;
            ++cleanup;
        }
    }
    // copy coefficients/RHS (:437-441) and the basis, shard by shard; scales of the re-pricing from
    // the ORIGINAL objective row (the basis columns are exact unit vectors, so the value the
    // sequential loop of :444-451 reads for row i is the original one)
    for (size_t g = 0; g < nsh; ++g) {
        Shard &as = art->shards[g];
        Shard &ms = mn->shards[g];
        TP_CU(cudaSetDevice(ms.device));
        TP_CU(cudaStreamSynchronize(as.stream));
        if (ms.m_local > 0) {
            dim3 grid((unsigned)std::min<int64_t>((nv + 256) / 256, 64), (unsigned)std::min<int64_t>(ms.m_local, 32768));
            k_copy_art_to_main<<<grid, 256, 0, ms.stream>>>(as.tab, as.ld, (int)art_nv, ms.tab, ms.ld,
                                                            (int)nv, ms.m_local);
            TP_CU(cudaMemcpyAsync(ms.basis, as.basis, sizeof(int32_t) * ms.m_local, cudaMemcpyDeviceToDevice, ms.stream));
            TP_CU(cudaMalloc(&d_scales[g], sizeof(double) * ms.m_local));
            k_reprice_scales<<<(unsigned)((ms.m_local + 255) / 256), 256, 0, ms.stream>>>(
                ms.tab + (int64_t)ms.m_local * ms.ld, ms.basis, ms.m_local, (int)nv, d_scales[g]);
        }
        TP_CU(cudaStreamSynchronize(ms.stream));
        TP_CU(cudaGetLastError());
    }
    // the objective row travels through the shards in rank order
    for (size_t g = 0; g < nsh; ++g) {
        Shard &ms = mn->shards[g];
        TP_CU(cudaSetDevice(ms.device));
        double *obj_g = ms.tab + (int64_t)ms.m_local * ms.ld;
        if (g > 0) {
            Shard &pv = mn->shards[g - 1];
            TP_CU(cudaMemcpyPeer(obj_g, ms.device, pv.tab + (int64_t)pv.m_local * pv.ld, pv.device,
                                 sizeof(double) * C));
        }
        if (ms.m_local > 0)
            k_reprice<<<(unsigned)((nv + 1 + 127) / 128), 128, 0, ms.stream>>>(ms.tab, ms.ld, ms.m_local,
                                                                              (int)nv, d_scales[g]);
        TP_CU(cudaStreamSynchronize(ms.stream));
        TP_CU(cudaGetLastError());
    }
    for (size_t g = 0; g + 1 < nsh; ++g) {                 // every replica gets the re-priced row
        Shard &ms = mn->shards[g];
        Shard &last = mn->shards[nsh - 1];
        TP_CU(cudaMemcpyPeer(ms.tab + (int64_t)ms.m_local * ms.ld, ms.device,
                             last.tab + (int64_t)last.m_local * last.ld, last.device, sizeof(double) * C));
    }
    TP_CU(cudaSetDevice(mn->shards[0].device));
    status = b200lp_iterate(mn, 0, &r2, nullptr, nullptr);
    if (status >= 0) {
        if (opts && opts->writeback_full) {
            if ((rc = b200lp_download(art, art_tab, ld_art, art_basis))) return finish(rc);
            if ((rc = b200lp_download(mn, main_tab, ld, main_basis))) return finish(rc);
        } else {
            std::vector<double> rhs((size_t)R), obj((size_t)C);
            if ((rc = b200lp_download_solution(mn, rhs.data(), obj.data(), main_basis))) return finish(rc);
            for (int64_t i = 0; i < R; ++i) main_tab[i * ld + (C - 1)] = rhs[(size_t)i];
            std::memcpy(main_tab + (R - 1) * ld, obj.data(), sizeof(double) * C);
        }
    }
#undef TP_CU
    return finish(status);
}

} // extern "C"
