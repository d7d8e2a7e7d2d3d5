/*
 * b200lp.h -- C ABI of libb200lp.so, the B200-native dense tableau simplex backend that
 * plugs in behind neil-lindquist/linear-programming's `*solver*` hook.
 *
 * Boundary (reference @ 7fe5c78, all paths under /root/reference):
 *   src/solver.lisp:39-56   `*solver*` / `solve-problem` -- the hook a backend function serves.
 *   src/simplex.lisp:48-58  `tableau` struct            -- the data this ABI receives:
 *       matrix  row-major (m+1) x (var-count+1); last column = right-hand side, last row =
 *       objective row (:74-78); basis-columns[i] = column basic in row i (:54).
 *   src/simplex.lisp:399-461 `n-solve-tableau`          -- what b200lp_solve* replace.
 *   src/conditions.lisp:43-77 solver-error hierarchy    -- what the status codes map onto.
 * The reference is pure Common Lisp and has no FFI of its own; INTEGRATION.md shows the CFFI
 * binding (and the ctypes binding this repo's tests use) for every entry point below.
 *
 * Conventions: plain pointers and sizes only; caller owns every buffer; every function returns
 * B200LP_OK (0) or a status/err code and never throws or aborts across the ABI; all calls on one
 * handle must come from one thread at a time (the library serialises internally).  All tableau
 * arithmetic is IEEE fp64 with a separately rounded product and difference (no FMA), true
 * division for the pivot-row scale, lowest-index tie-breaks -- bit-identical to the reference's
 * double-float arithmetic (src/simplex.lisp:344-357).
 */
#ifndef B200LP_H
#define B200LP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LP_VERSION 200 /* 0.2.0 */
#define B200LP_MAX_DEVICES 8

/* ---- status codes --------------------------------------------------------------------------
 * >= 0: solver outcomes (src/conditions.lisp); < 0: argument / CUDA / NCCL failures, which the
 * Lisp shim signals as plain `solver-error`. */
enum {
    B200LP_OK = 0,                /* optimal tableau reached (n-solve-tableau returned)          */
    B200LP_UNBOUNDED = 1,         /* src/simplex.lisp:458-459 -> unbounded-problem-error         */
    B200LP_INFEASIBLE = 2,        /* src/simplex.lisp:405-407 -> infeasible-problem-error        */
    B200LP_ITERATION_LIMIT = 3,   /* build extension (reference has no cap): opts.max_iters hit  */
    B200LP_ARTIFICIAL_STUCK = 4,  /* src/simplex.lisp:432-433 (plain `error`): an artificial variable
                                     is still basic and no column can replace it (B200LP_FEAS_REFERENCE only) */
    B200LP_ARTIFICIAL_NONZERO = 5,/* src/simplex.lisp:423-424 (plain `error`): an artificial variable
                                     is still basic at a non-zero level                           */
    B200LP_ERR_INVALID_ARG = -1,
    B200LP_ERR_CUDA = -2,
    B200LP_ERR_NCCL = -3,
    B200LP_ERR_NO_DEVICE = -4,
    B200LP_ERR_OUT_OF_MEMORY = -5,
    B200LP_ERR_INTERNAL = -6,
    B200LP_ERR_PEER_TIMEOUT = -7  /* sharded: a peer GPU's candidate row never arrived            */
};

/* How the two-phase transition (src/simplex.lisp:402-434) judges "zero".  The reference keeps
 * exact rationals exact, so its ABSOLUTE tolerance tol*eps (:405-406) and its exact `/= 0`
 * tests (:423, :427) only ever meet floating-point residue when the user's input is floating
 * point.  This backend coerces everything to fp64, so by default (B200LP_FEAS_SCALED) it scales
 * the tolerance by s = max(1, |initial phase-1 objective|): infeasible iff |obj| > tol*eps*s; a
 * basic artificial must have |RHS| <= tol*eps*s; it is replaced by the non-basic column with the
 * largest |entry| > (tol/2)*eps (first index on ties); a row with no such column is redundant and
 * keeps its artificial at level zero (no error).  B200LP_FEAS_REFERENCE applies the reference's
 * tests literally.  oracle/simplex_oracle.c implements both, and the parity tests cover both. */
enum { B200LP_FEAS_SCALED = 0, B200LP_FEAS_REFERENCE = 1 };

enum { B200LP_RULE_REFERENCE = 0, /* Dantzig, first index on ties: src/simplex.lisp:362-389      */
       B200LP_RULE_BLAND = 1 };   /* build extension: anti-cycling                 */

/* ---- options (every field zero = defaults) ------------------------------------------------- */
typedef struct b200lp_opts {
    double  fp_tolerance_factor; /* `:fp-tolerance` (src/simplex.lisp:511); 0 -> 1024              */
    int32_t pivot_rule;          /* B200LP_RULE_*                                                  */
    int32_t writeback_full;      /* one-shot calls: copy the whole solved tableau back to `tab`    */
    int64_t max_iters;           /* 0 = unlimited, like the reference                              */
    int32_t ndev;                /* 0/1 = one GPU; N = row-block shard over devices[0..N) in-process*/
    int32_t devices[B200LP_MAX_DEVICES]; /* CUDA ordinals; with ndev == 0 devices[0] is used        */
    int32_t trace_capacity;      /* record (entering col, leaving row) of the first N pivots        */
    int32_t poll_interval;       /* pivots enqueued between host polls of the status word; 0 = auto */
    int32_t time_kernels;        /* record CUDA events around every pivot-update launch             */
    int32_t pivot_variant;       /* tuning knob for the rank-1 update kernel; 0 = default           */
    int32_t feas_mode;           /* B200LP_FEAS_* (two-phase calls)                                 */
    int32_t reserved[5];
} b200lp_opts;

/* ---- result / telemetry -------------------------------------------------------------------- */
typedef struct b200lp_result {
    int32_t status;              /* same value the call returned                                    */
    int32_t n_devices;
    int64_t iterations;          /* pivots performed by this call (phase 2 for two-phase)           */
    int64_t iterations_phase1;   /* two-phase only                                                  */
    int64_t iterations_cleanup;  /* two-phase only: zero-level artificial pivots (:419-434)         */
    double  objective;           /* matrix[m, var-count] (src/simplex.lisp:74-78)                   */
    double  ms_total;            /* wall time inside the call                                       */
    double  ms_h2d;              /* host->device copy of the tableau                                */
    double  ms_solve;            /* device time of the iteration loop (CUDA events)                 */
    double  ms_d2h;              /* device->host copy of the results                                */
    double  ms_pivot_kernel;     /* sum of pivot-update launch durations (time_kernels only)        */
    int64_t pivot_kernel_launches;
    int64_t kernel_launches;     /* every kernel this call launched                                 */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t bytes_per_pivot;     /* algorithmic: 16 * R * C, per device: 16*R_local*C   */
    int32_t trace_len;           /* entries valid in the trace buffers                              */
    int32_t exchange_mode;       /* sharded candidates: 0 one shard, 1 NCCL all-gather, 2 peer-mapped */
    double  ms_look_kernel;      /* sum of lookahead-kernel durations (time_kernels only)           */
    double  ms_exchange;         /* sum of candidate-exchange durations (time_kernels, sharded)     */
    int64_t look_kernel_launches;
    /* how the loop ran and where the decision chain spent its time (persistent loop, summed over
     * pivots, %globaltimer on look CTA 0) */
    int32_t loop_mode;           /* 1 = one k_iter launch per pivot, 2 = one persistent cooperative kernel */
    int32_t look_ctas;
    double  ms_look_wait;        /* waiting for the tile CTAs to finish update(k-2)                 */
    double  ms_look_ratio;       /* phase A: pivot-column gather + ratio test                       */
    double  ms_look_push;        /* sharded: scale own candidate row + push to every rank           */
    double  ms_look_peer_wait;   /* sharded: waiting for every rank's candidate                     */
    double  ms_look_row;         /* phase B: pivot row / element, objective row, next entering col  */
    double  sm_clock_mhz;        /* SM clock seen by the look role (clock64 vs %globaltimer)        */
    double  ms_look_dbg[8];      /* finer split of the two phases (dev aid; see persist.cuh)        */
    int64_t redundant_rows;      /* two-phase, B200LP_FEAS_SCALED: rows left with a zero-level artificial */
    int32_t look_cluster;        /* k_persist: thread-block cluster size of the look grid (0 = none) */
    int32_t reserved_r;
} b200lp_result;

/* ---- one-shot calls: what the `*solver*` backend function uses ------------------------------
 * Replace (n-solve-tableau tableau), src/simplex.lisp:453-461.
 *   tab    host, row-major, R rows x C columns, `ld` doubles per row (ld >= C).
 *   basis  host, R-1 entries.   is_max: 1 for `max` problems, 0 for `min` (:365).
 * On return (status >= 0) column C-1, row R-1 and basis[] hold the solved values (everything
 * tableau-objective-value / tableau-variable / tableau-reduced-cost read, :74-120); the whole
 * matrix only if opts->writeback_full.  trace_j/trace_r may be NULL. */
int b200lp_solve(const b200lp_opts *opts, double *tab, int64_t R, int64_t C, int64_t ld,
                 int32_t *basis, int32_t is_max, b200lp_result *out,
                 int32_t *trace_j, int32_t *trace_r);

/* Replace (n-solve-tableau (list art-tableau main-tableau)), src/simplex.lisp:402-452:
 * phase 1 on the artificial tableau (always a `min` problem, :317-319), feasibility check,
 * zero-level artificial clean-up, coefficient copy, objective re-pricing, phase 2 on `main_tab`.
 * Both tableaus have R rows.  On return art_tab/art_basis hold the solved phase-1 tableau only
 * if opts->writeback_full; main_tab/main_basis as for b200lp_solve.  With opts->ndev > 1 both
 * tableaus are row-block sharded like b200lp_solve's; the transition's order-dependent step (the
 * objective re-pricing over all rows in order) hands the objective row from shard to shard in
 * rank order, so the result is bit-identical to the one-GPU run. */
int b200lp_solve_two_phase(const b200lp_opts *opts,
                           double *art_tab, int64_t C_art, int64_t ld_art, int32_t *art_basis,
                           double *main_tab, int64_t R, int64_t C, int64_t ld, int32_t *main_basis,
                           int32_t is_max, b200lp_result *out);

/* ---- resident-tableau handle: the same path step by step ------------------------------------
 * Used by the parity tests (one reference function per call), by bench.py (tableau already in
 * HBM when the clock starts) and by a branch-and-bound driver that keeps the tableau on device. */
typedef struct b200lp_solver b200lp_solver;

int b200lp_create(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max,
                  b200lp_solver **out);
void b200lp_destroy(b200lp_solver *s);

/* copy-tableau in (src/simplex.lisp:61-71 plays this role on the host) */
int b200lp_upload(b200lp_solver *s, const double *tab, int64_t ld, const int32_t *basis);
/* full matrix + basis out; either pointer may be NULL */
int b200lp_download(b200lp_solver *s, double *tab, int64_t ld, int32_t *basis);
/* only what the solution accessors read: rhs[R] (column C-1), obj_row[C] (row R-1), basis[R-1] */
int b200lp_download_solution(b200lp_solver *s, double *rhs, double *obj_row, int32_t *basis);

/* find-entering-column, src/simplex.lisp:362-379.  *col = -1 when none (tableau optimal). */
int b200lp_find_entering_column(b200lp_solver *s, int64_t *col);
/* find-pivoting-row, src/simplex.lisp:382-389.  *row = -1 when none (unbounded direction). */
int b200lp_find_pivoting_row(b200lp_solver *s, int64_t entering_col, int64_t *row);
/* n-pivot-row, src/simplex.lisp:337-359. */
int b200lp_pivot(b200lp_solver *s, int64_t entering_col, int64_t changing_row);
/* n-solve-tableau loop (:455-460) for at most max_iters more pivots (0 = opts.max_iters /
 * unlimited).  Returns B200LP_OK / UNBOUNDED / ITERATION_LIMIT. */
int b200lp_iterate(b200lp_solver *s, int64_t max_iters, b200lp_result *out,
                   int32_t *trace_j, int32_t *trace_r);

/* Turn per-launch CUDA-event timing (opts.time_kernels) on or off for later b200lp_iterate calls.
 * Timing events between launches serialise consecutive iterations, so throughput runs keep it off. */
int b200lp_set_time_kernels(b200lp_solver *s, int32_t on);

/* ---- multi-process row-block sharding (one process per GPU, NCCL over NVLink) ---------------
 * Rank g owns constraint rows [row_begin, row_end) plus a replica of the objective row; the
 * per-iteration exchange is one all-gather of each rank's candidate pivot row.  `unique_id` is
 * 128 bytes from b200lp_comm_unique_id() on rank 0, distributed by the caller (torch.distributed,
 * MPI, a file ...).  upload/download then take the LOCAL block: (row_end-row_begin)+1 rows, the
 * last one being the objective row; basis has row_end-row_begin entries. */
int b200lp_comm_unique_id(void *unique_id_128);
int b200lp_create_sharded(const b200lp_opts *opts, int64_t R, int64_t C, int32_t is_max,
                          int32_t rank, int32_t nranks, const void *unique_id_128,
                          b200lp_solver **out);
int b200lp_shard_rows(const b200lp_solver *s, int64_t *row_begin, int64_t *row_end);
/* the partition rule itself (host arithmetic, no GPU): contiguous blocks of ceil(m/nranks) */
void b200lp_partition(int64_t m, int32_t nranks, int32_t rank, int64_t *row_begin, int64_t *row_end);

/* The one-shot calls keep up to four idle single-GPU handles (<= 64 MB of tableau each) so that
 * runs of small solves -- branch and bound, test suites -- skip the ~3 ms of allocation per call.
 * b200lp_shutdown() frees them; optional (they are also reclaimed at process exit). */
void b200lp_shutdown(void);

/* ---- misc ----------------------------------------------------------------------------------- */
const char *b200lp_strerror(int code);
int b200lp_version(void);
/* sizeof(b200lp_opts), sizeof(b200lp_result): a binding that declares the structs by hand (the
 * CFFI shim, ctypes) asserts its own sizes against these when it loads */
void b200lp_abi_sizes(int64_t *opts_size, int64_t *result_size);
int b200lp_device_count(void);            /* CUDA devices visible; 0 when there is none          */
const char *b200lp_last_error(void);      /* text of the last CUDA/NCCL failure on this thread   */
/* thresholds actually used, for parity tests: (tol/8)eps, (tol/2)eps, tol*eps with
 * eps = CL double-float-epsilon = 0x1.0000000000001p-53 (src/utils.lisp:92,107) */
void b200lp_thresholds(double fp_tolerance_factor, double *enter, double *pivot, double *feas);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_H */
