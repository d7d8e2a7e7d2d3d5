/*
 * simplex_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar fp64 CPU restatement of the reference's dense tableau simplex hot path
 * (neil-lindquist/linear-programming @ 7fe5c78, src/simplex.lisp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libb200lp.so) never links or calls it.
 *
 * Parity pin: the reference (Common Lisp) cannot run in this image (no Lisp
 * implementation); this restatement is pinned by every golden tableau the
 * reference's own tests hold for the path (t/simplex.lisp:60-275, README.md:58-62,
 * t/integration.lisp:32-58) via tests/test_oracle_golden.py, which also cross-checks
 * it against the exact-rational restatement in oracle/exact.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  -ffp-contract=off
 * matters: the reference rounds the product and the difference separately
 * (src/simplex.lisp:357, `decf (aref ...) (* scale ...)`), so no FMA may be formed.
 *
 * Layout (src/simplex.lisp:48-58, 214-287): row-major R x C doubles with leading
 * dimension ld >= C; rows 0..m-1 constraints, row m = R-1 the objective row; column
 * C-1 the right-hand side; var_count = C-1; basis[i] = column basic in row i.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* CL double-float-epsilon as SBCL defines it: 2^-53 * (1 + 2^-52)
 * (src/utils.lisp:92,107 multiply `factor` by this constant). */
#define CL_DOUBLE_FLOAT_EPSILON 0x1.0000000000001p-53

enum {
    ORACLE_OPTIMAL = 0,
    ORACLE_UNBOUNDED = 1,      /* src/simplex.lisp:458-459 unbounded-problem-error */
    ORACLE_INFEASIBLE = 2,     /* src/simplex.lisp:405-407 infeasible-problem-error */
    ORACLE_ITERATION_LIMIT = 3,/* build extension: the reference has no cap */
    ORACLE_ARTIFICIAL_STUCK = 4,/* src/simplex.lisp:432-433 plain `error`: cannot be replaced */
    ORACLE_ARTIFICIAL_NONZERO = 5/* src/simplex.lisp:423-424 plain `error`: still non-zero */
};

/* How the two-phase transition judges "zero" (build extension):
 * FEAS_REFERENCE: exactly the reference -- |obj| <= tol*eps absolute (:405-406), artificial
 *   rows need RHS == 0 and a replacement column with an exactly non-zero entry (:419-434).
 *   Right for exact-rational input, which the reference keeps exact; on fp64 it rejects
 *   feasible problems whose phase-1 objective carries rounding residue.
 * FEAS_SCALED (the backend's default): the same tests with the tolerance scaled by the size
 *   of the numbers that cancelled, s = max(1, |initial phase-1 objective|): |obj| <= tol*eps*s,
 *   artificial rows need |RHS| <= tol*eps*s, the replacement column is the non-basic column
 *   with the LARGEST |entry| > (tol/2)*eps (first index on ties), and a row with no such column
 *   is redundant and keeps its artificial at level zero instead of raising. */
enum { FEAS_SCALED = 0, FEAS_REFERENCE = 1 };

enum { RULE_REFERENCE = 0, RULE_BLAND = 1 };

double oracle_cl_epsilon(void) { return CL_DOUBLE_FLOAT_EPSILON; }

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Threads for the OpenMP-over-rows pivot (CPU baseline only; results do not depend on it).
 * torchrun exports OMP_NUM_THREADS=1, so the bench sets the count explicitly. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* find-entering-column, src/simplex.lisp:362-379.
 * max problem: first argmin of the objective row over [0, var_count); accept iff
 * value < 0 - (tol/8)*eps.  min problem: first argmax; accept iff > 0 + (tol/8)*eps.
 * Iterate's `finding .. minimizing` keeps the first extremum (strict compare).
 * rule 1 (Bland, build extension): lowest index passing the same threshold test. */
int64_t oracle_find_entering_column(const double *tab, int64_t R, int64_t C, int64_t ld,
                                    int is_max, double tol, int rule)
{
    const double *obj = tab + (R - 1) * ld;
    const int64_t nv = C - 1;
    const double thr = (tol / 8.0) * CL_DOUBLE_FLOAT_EPSILON;
    if (nv <= 0) return -1;
    if (rule == RULE_BLAND) {
        for (int64_t i = 0; i < nv; ++i) {
            if (is_max ? (obj[i] < 0.0 - thr) : (obj[i] > 0.0 + thr)) return i;
        }
        return -1;
    }
    int64_t best = 0;
    if (is_max) {
        for (int64_t i = 1; i < nv; ++i) if (obj[i] < obj[best]) best = i;
        return (obj[best] < 0.0 - thr) ? best : -1;
    } else {
        for (int64_t i = 1; i < nv; ++i) if (obj[i] > obj[best]) best = i;
        return (obj[best] > 0.0 + thr) ? best : -1;
    }
}

/* find-pivoting-row, src/simplex.lisp:382-389.
 * Over rows [0, m) with a[i,j] > (tol/2)*eps: first argmin of rhs_i / a[i,j].
 * No lower clamp on the ratio.  -1 when no row is eligible (=> unbounded).
 * rule 1 (Bland): among exact ratio ties the row with the smallest basis index. */
int64_t oracle_find_pivoting_row(const double *tab, int64_t R, int64_t C, int64_t ld,
                                 const int32_t *basis, int64_t j, double tol, int rule)
{
    const int64_t m = R - 1;
    const double thr = (tol / 2.0) * CL_DOUBLE_FLOAT_EPSILON;
    int64_t best_row = -1;
    double best_q = 0.0;
    for (int64_t i = 0; i < m; ++i) {
        const double a = tab[i * ld + j];
        if (0.0 + thr < a) {
            const double q = tab[i * ld + (C - 1)] / a;
            if (best_row < 0 || q < best_q ||
                (rule == RULE_BLAND && q == best_q && basis[i] < basis[best_row])) {
                best_q = q;
                best_row = i;
            }
        }
    }
    return best_row;
}

/* n-pivot-row, src/simplex.lisp:337-359.
 * (1) row p <- row p / a[p,j] over all C columns (true division);
 * (2) every other row r in [0,R) including the objective row: scale = a[r,j] read
 *     before the row changes, then a[r,c] -= scale * a[p,c] (rounded product, then
 *     rounded difference); (3) basis[p] = j.  No zero skipping.
 * Rows are independent once row p is scaled, so OpenMP over rows changes nothing
 * in the values. */
void oracle_pivot(double *tab, int64_t R, int64_t C, int64_t ld, int32_t *basis,
                  int64_t j, int64_t p, int parallel)
{
    double *prow = tab + p * ld;
    const double row_scale = prow[j];
    for (int64_t c = 0; c < C; ++c) prow[c] = prow[c] / row_scale;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (parallel)
#endif
    for (int64_t r = 0; r < R; ++r) {
        if (r == p) continue;
        double *row = tab + r * ld;
        const double scale = row[j];
        for (int64_t c = 0; c < C; ++c) {
            const double prod = scale * prow[c];
            row[c] = row[c] - prod;
        }
    }
    (void)parallel;
    if (basis) basis[p] = (int32_t)j;
}

/* n-solve-tableau, single-tableau branch, src/simplex.lisp:453-461.
 * trace_j/trace_r (optional, capacity trace_cap) record (entering col, leaving row)
 * per pivot.  max_iters == 0 means unlimited, as in the reference. */
int oracle_solve(double *tab, int64_t R, int64_t C, int64_t ld, int32_t *basis,
                 int is_max, double tol, int rule, int64_t max_iters, int parallel,
                 int64_t *iters_out, int32_t *trace_j, int32_t *trace_r, int64_t trace_cap)
{
    int64_t it = 0;
    int status = ORACLE_OPTIMAL;
    for (;;) {
        const int64_t j = oracle_find_entering_column(tab, R, C, ld, is_max, tol, rule);
        if (j < 0) break;
        if (max_iters > 0 && it >= max_iters) { status = ORACLE_ITERATION_LIMIT; break; }
        const int64_t p = oracle_find_pivoting_row(tab, R, C, ld, basis, j, tol, rule);
        if (p < 0) { status = ORACLE_UNBOUNDED; break; }
        if (it < trace_cap) {
            if (trace_j) trace_j[it] = (int32_t)j;
            if (trace_r) trace_r[it] = (int32_t)p;
        }
        oracle_pivot(tab, R, C, ld, basis, j, p, parallel);
        ++it;
    }
    if (iters_out) *iters_out = it;
    return status;
}

/* n-solve-tableau, list (two-phase) branch, src/simplex.lisp:402-452.
 * art: R x C_art (ld_art), its instance problem is `min` (src/simplex.lisp:317-319);
 * main: R x C (ld), min/max per is_max.  On success the main tableau holds the
 * phase-2 optimum and main_basis its basis.  iters_out[0] = phase-1 pivots,
 * iters_out[1] = artificial clean-up pivots, iters_out[2] = phase-2 pivots,
 * iters_out[3] = redundant rows left with a zero-level artificial (FEAS_SCALED only). */
int oracle_solve_two_phase(double *art, int64_t C_art, int64_t ld_art, int32_t *art_basis,
                           double *mtab, int64_t R, int64_t C, int64_t ld, int32_t *main_basis,
                           int is_max, double tol, int rule, int64_t max_iters, int parallel,
                           int feas_mode, int64_t *iters_out)
{
    const int64_t m = R - 1;
    const int64_t num_vars = C - 1;         /* main var-count */
    const int64_t num_art_vars = C_art - 1; /* art var-count  */
    int64_t it1 = 0, it_fix = 0, it2 = 0, redundant = 0;
    if (iters_out) iters_out[0] = iters_out[1] = iters_out[2] = iters_out[3] = 0;
    const double obj0 = art[m * ld_art + num_art_vars];
    const double scale = (feas_mode == FEAS_REFERENCE) ? 1.0 : fmax(1.0, fabs(obj0));
    const double thr_feas = tol * CL_DOUBLE_FLOAT_EPSILON * scale;
    const double thr_pivot = (tol / 2.0) * CL_DOUBLE_FLOAT_EPSILON;

    int st = oracle_solve(art, R, C_art, ld_art, art_basis, /*is_max=*/0, tol, rule,
                          max_iters, parallel, &it1, NULL, NULL, 0);
    if (iters_out) iters_out[0] = it1;
    if (st != ORACLE_OPTIMAL) return st;

    /* (unless (fp= 0 obj tol) (error 'infeasible-problem-error)) :405-407 */
    const double art_obj = art[m * ld_art + num_art_vars];
    if (!(fabs(0.0 - art_obj) <= thr_feas)) return ORACLE_INFEASIBLE;

    /* drive zero-level artificials out of the basis :419-434 */
    for (int64_t i = 0; i < m; ++i) {
        if (art_basis[i] >= num_vars) {
            const double rhs = art[i * ld_art + num_art_vars];
            if (feas_mode == FEAS_REFERENCE ? (rhs != 0.0) : !(fabs(rhs) <= thr_feas))
                return ORACLE_ARTIFICIAL_NONZERO;
            int64_t new_col = -1;
            double best = 0.0;
            for (int64_t j = 0; j < num_vars; ++j) {
                const double a = art[i * ld_art + j];
                const int cand = (feas_mode == FEAS_REFERENCE) ? (a != 0.0)
                                                               : (fabs(a) > thr_pivot && fabs(a) > best);
                if (cand) {
                    int in_basis = 0;
                    for (int64_t k = 0; k < m; ++k)
                        if (art_basis[k] == j) { in_basis = 1; break; }
                    if (!in_basis) {
                        new_col = j;
                        if (feas_mode == FEAS_REFERENCE) break;
                        best = fabs(a);
                    }
                }
            }
            if (new_col < 0) {
                if (feas_mode == FEAS_REFERENCE) return ORACLE_ARTIFICIAL_STUCK;
                ++redundant;                 /* the row is a combination of the others */
                continue;
            }
            oracle_pivot(art, R, C_art, ld_art, art_basis, new_col, i, parallel);
            ++it_fix;
        }
    }
    if (iters_out) { iters_out[1] = it_fix; iters_out[3] = redundant; }

    /* copy coefficients and RHS of the constraint rows :437-441 */
    for (int64_t r = 0; r < m; ++r) {
        for (int64_t c = 0; c < num_vars; ++c) mtab[r * ld + c] = art[r * ld_art + c];
        mtab[r * ld + num_vars] = art[r * ld_art + num_art_vars];
    }
    /* basis + re-price the objective row :444-451 (a redundant row's artificial has no column
     * in the main tableau: nothing to price out) */
    double *obj = mtab + m * ld;
    for (int64_t i = 0; i < m; ++i) {
        const int32_t bc = art_basis[i];
        main_basis[i] = bc;
        if (bc >= num_vars) continue;
        const double scale_i = obj[bc];
        if (scale_i != 0.0) {
            const double *row = mtab + i * ld;
            for (int64_t c = 0; c <= num_vars; ++c) {
                const double prod = scale_i * row[c];
                obj[c] = obj[c] - prod;
            }
        }
    }
    st = oracle_solve(mtab, R, C, ld, main_basis, is_max, tol, rule, max_iters, parallel,
                      &it2, NULL, NULL, 0);
    if (iters_out) iters_out[2] = it2;
    return st;
}

/* n-solve-tableau with a cycle probe (config 5's question: does the reference rule CYCLE on a
 * degenerate LP, or merely stall?).  The reference has no anti-cycling (src/simplex.lisp:455-460),
 * so a revisited basis -- the same column basic in the same row, for every row -- under a
 * deterministic rule is a proven cycle.  The basis vector is hashed incrementally (one term per
 * row, replaced on every pivot); hashes live in an open-addressing table sized for max_iters.
 * out[0] = pivots done, out[1] = degenerate pivots (objective value unchanged), out[2] = pivot
 * index at which a basis was first revisited (-1: never), out[3] = pivot index of its first visit. */
static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

int oracle_solve_cycle_probe(double *tab, int64_t R, int64_t C, int64_t ld, int32_t *basis,
                             int is_max, double tol, int rule, int64_t max_iters, int parallel,
                             int64_t *out)
{
    const int64_t m = R - 1;
    int64_t cap = 1;
    while (cap < 4 * (max_iters + 2)) cap <<= 1;
    uint64_t *keys = (uint64_t *)calloc((size_t)cap, sizeof(uint64_t));
    int64_t *when = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
    if (!keys || !when) { free(keys); free(when); return -1; }
    uint64_t h = 0;
    for (int64_t i = 0; i < m; ++i) h += mix64(((uint64_t)i << 32) | (uint32_t)basis[i]);
    int64_t it = 0, degenerate = 0, revisit = -1, first = -1;
    int status = ORACLE_OPTIMAL;
    for (;;) {
        /* record / look up the current basis */
        if (revisit < 0) {
            const uint64_t key = h | 1ULL;               /* 0 marks an empty slot */
            int64_t pos = (int64_t)(mix64(key) & (uint64_t)(cap - 1));
            while (keys[pos] != 0 && keys[pos] != key) pos = (pos + 1) & (cap - 1);
            if (keys[pos] == key) { revisit = it; first = when[pos]; }
            else { keys[pos] = key; when[pos] = it; }
        }
        const int64_t j = oracle_find_entering_column(tab, R, C, ld, is_max, tol, rule);
        if (j < 0) break;
        if (max_iters > 0 && it >= max_iters) { status = ORACLE_ITERATION_LIMIT; break; }
        const int64_t p = oracle_find_pivoting_row(tab, R, C, ld, basis, j, tol, rule);
        if (p < 0) { status = ORACLE_UNBOUNDED; break; }
        const double before = tab[m * ld + (C - 1)];
        h -= mix64(((uint64_t)p << 32) | (uint32_t)basis[p]);
        oracle_pivot(tab, R, C, ld, basis, j, p, parallel);
        h += mix64(((uint64_t)p << 32) | (uint32_t)basis[p]);
        if (tab[m * ld + (C - 1)] == before) ++degenerate;
        ++it;
    }
    free(keys); free(when);
    if (out) { out[0] = it; out[1] = degenerate; out[2] = revisit; out[3] = first; }
    return status;
}
