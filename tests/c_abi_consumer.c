/* A plain-C consumer of libb200lp.so: what any FFI host (CFFI, cgo, JNI ...) does.  Solves the
 * README LP (t/simplex.lisp:60-72, 170-194) through b200lp_solve.  Exit code 0 = solved and
 * correct on a GPU box, or refused with B200LP_ERR_NO_DEVICE on a box without a GPU (there is
 * no CPU fallback); anything else is a failure.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>
#include "b200lp.h"

int main(void)
{
    double tab[3 * 6] = {2, 1, 0, 1, 0, 8,
                         0, 1, 1, 0, 1, 7,
                         -1, -4, -3, 0, 0, 0};
    int32_t basis[2] = {3, 4};
    int32_t tj[8], tr[8];
    b200lp_opts opts;
    b200lp_result res;
    double e, p, f;
    int64_t b0, b1;
    int rc;

    memset(&opts, 0, sizeof opts);                 /* every zero field = default */
    opts.trace_capacity = 8;
    if (b200lp_version() != B200LP_VERSION) return 10;
    b200lp_thresholds(1024.0, &e, &p, &f);
    if (!(e > 0 && p == 4 * e && f == 8 * e)) return 11;
    b200lp_partition(10, 3, 1, &b0, &b1);
    if (b0 != 4 || b1 != 8) return 12;

    rc = b200lp_solve(&opts, tab, 3, 6, 6, basis, 1, &res, tj, tr);
    if (b200lp_device_count() == 0) {
        printf("no device: rc=%d (%s)\n", rc, b200lp_strerror(rc));
        return rc == B200LP_ERR_NO_DEVICE ? 0 : 13;
    }
    if (rc != B200LP_OK) { printf("rc=%d %s %s\n", rc, b200lp_strerror(rc), b200lp_last_error()); return 14; }
    if (res.iterations != 2 || res.objective != 28.5) return 15;
    if (basis[0] != 0 || basis[1] != 1) return 16;
    if (tab[5] != 0.5 || tab[11] != 7.0 || tab[17] != 28.5) return 17;   /* RHS column */
    if (res.trace_len != 2 || tj[0] != 1 || tr[0] != 1 || tj[1] != 0 || tr[1] != 0) return 18;
    b200lp_shutdown();
    printf("solved: objective %.17g in %lld pivots, %lld kernel launches\n", res.objective,
           (long long)res.iterations, (long long)res.kernel_launches);
    return 0;
}
