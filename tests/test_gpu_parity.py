"""Parity of the CUDA path (through the C ABI, ctypes) against the oracle -- bit-exact.

Reads like the reference's t/simplex.lisp: pivot-row, basic-problem, equality-constraint,
leq-constraint, unsolvable-problems, plus randomized differential tests the reference lacks.
All tests need a real B200 (-m gpu).  The oracle is only the checker here.
"""
import numpy as np
import pytest

from golden import reference_goldens as G
from lp_checks import certify_optimal
from linear_programming_b200 import _ffi, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["1", "0"], ids=["one-cta-on", "one-cta-off"])
def small_path(request, monkeypatch):
    """Tableaus that fit one CTA's shared memory are solved by k_small (one launch per solve);
    B200LP_SMALL=0 sends them down the big path (k_persist / k_iter) instead.  Both must equal the
    oracle bit for bit."""
    monkeypatch.setenv("B200LP_SMALL", request.param)
    return request.param == "1"


def f64(rows):
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float64)


def i32(b):
    return np.array(b, dtype=np.int32)


def random_tableau(m, n, seed, signed=False):
    rng = np.random.default_rng(seed)
    A = rng.random((m, n)) - (0.25 if signed else 0.0)
    b = rng.uniform(n / 8.0, 3.0 * n / 8.0, m)
    c = rng.random(n)
    return synthetic.tableau_from_lp(A, b, c)


# ------------------------------------------------------------------ reference goldens
def test_pivot_row_golden():
    """t/simplex.lisp:135-159"""
    g = G.SINGLE_PIVOT
    tab, basis = f64(g["initial"]["matrix"]), i32(g["initial"]["basis"])
    with _ffi.DeviceTableau(*tab.shape) as d:
        d.upload(tab, basis)
        d.pivot(g["col"], g["row"])
        out, ob = d.download()
    assert np.array_equal(out, f64(g["matrix"])) and ob.tolist() == g["basis"]
    assert out[-1, -1] == 4.0


def test_basic_problem_golden(small_path):
    """t/simplex.lisp:170-194, README.md:58-62"""
    g = G.BASIC_SOLVED
    tab, basis = f64(g["initial"]["matrix"]), i32(g["initial"]["basis"])
    st, res, trace = _ffi.solve(tab, basis, True, _ffi.make_opts(writeback_full=True, trace_capacity=8))
    assert st == _ffi.OK and res.iterations == g["pivots"] and trace == g["trace"]
    assert np.array_equal(tab, f64(g["matrix"])) and basis.tolist() == g["basis"]
    assert res.objective == 28.5
    assert res.loop_mode == (3 if small_path else 2) and (res.kernel_launches == 1 or not small_path)


def test_basic_problem_partial_writeback_touches_only_solution_cells():
    g = G.BASIC_SOLVED
    tab, basis = f64(g["initial"]["matrix"]), i32(g["initial"]["basis"])
    before = tab.copy()
    st, res, _ = _ffi.solve(tab, basis, True)
    want = f64(g["matrix"])
    assert st == _ffi.OK
    assert np.array_equal(tab[:, -1], want[:, -1]) and np.array_equal(tab[-1], want[-1])
    assert np.array_equal(tab[:-1, :-1], before[:-1, :-1])      # interior left alone
    assert basis.tolist() == g["basis"]


@pytest.mark.parametrize("g", [G.EQ_SOLVED, G.GEQ_SOLVED], ids=["eq", "geq"])
def test_two_phase_goldens(g, small_path):
    """t/simplex.lisp:196-275"""
    b = g["initial"]
    art, ab = f64(b["art_matrix"]), i32(b["art_basis"])
    main, mb = f64(b["main_matrix"]), i32(b["main_basis"])
    o_art, o_ab, o_main, o_mb = art.copy(), ab.copy(), main.copy(), mb.copy()
    st, res = _ffi.solve_two_phase(art, ab, main, mb, True, _ffi.make_opts(writeback_full=True))
    ost, oits = oracle.solve_two_phase(o_art, o_ab, o_main, o_mb, True)
    assert st == ost == _ffi.OK
    assert (res.iterations_phase1, res.iterations_cleanup, res.iterations) == oits == g["pivots"]
    assert ab.tolist() == g["art_basis"] and mb.tolist() == g["main_basis"]
    assert np.array_equal(art, o_art) and np.array_equal(main, o_main)        # bit-exact vs oracle
    np.testing.assert_allclose(main, f64(g["main_matrix"]), rtol=1e-14, atol=1e-14)
    assert abs(res.objective - float(g["objective"])) <= 1e-8 * float(g["objective"])


def test_unsolvable_problems(small_path):
    """t/simplex.lisp:277-289"""
    g = G.INFEASIBLE
    st, _ = _ffi.solve_two_phase(f64(g["art_matrix"]), i32(g["art_basis"]),
                                 f64(g["main_matrix"]), i32(g["main_basis"]), True)
    assert st == _ffi.INFEASIBLE
    g = G.UNBOUNDED
    st, _, _ = _ffi.solve(f64(g["matrix"]), i32(g["basis"]), True)
    assert st == _ffi.UNBOUNDED


def test_assembly_problem(small_path):
    """t/integration.lisp:32-58"""
    g = G.ASSEMBLY
    tab, basis = f64(g["matrix"]), i32(g["basis"])
    o_tab, o_basis = tab.copy(), basis.copy()
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(writeback_full=True))
    oracle.solve(o_tab, o_basis, True)
    assert st == _ffi.OK and res.iterations == g["pivots"]
    assert np.array_equal(tab, o_tab) and np.array_equal(basis, o_basis)
    lo, hi = g["bounds"]["revenue"]
    assert lo <= res.objective <= hi


# ------------------------------------------------------------- function-by-function parity
@pytest.mark.parametrize("m,n,signed", [(37, 53, False), (200, 300, False), (129, 64, True)])
def test_each_reference_function_matches_oracle(m, n, signed):
    tab, basis = random_tableau(m, n, seed=m * 1000 + n, signed=signed)
    o_tab, o_basis = tab.copy(), basis.copy()
    with _ffi.DeviceTableau(*tab.shape) as d:
        d.upload(tab, basis)
        for it in range(25):
            j = d.find_entering_column()
            oj = oracle.find_entering_column(o_tab, True)
            assert (j if j is not None else -1) == oj
            if j is None:
                break
            r = d.find_pivoting_row(j)
            orow = oracle.find_pivoting_row(o_tab, o_basis, j)
            assert (r if r is not None else -1) == orow
            if r is None:
                break
            d.pivot(j, r)
            oracle.pivot(o_tab, o_basis, j, r)
            g_tab, g_basis = d.download()
            assert np.array_equal(g_tab, o_tab), f"tableau differs after pivot {it}"
            assert np.array_equal(g_basis, o_basis)


@pytest.mark.parametrize("m,n,is_max", [(64, 96, True), (256, 512, True), (100, 40, True),
                                        (96, 160, False), (1, 3, True), (3, 1, True)])
def test_full_solve_bit_exact(m, n, is_max, small_path):
    tab, basis = random_tableau(m, n, seed=11 * m + n)
    if not is_max:
        tab[-1, :n] *= -1.0          # min (-c).x: the objective row (still -coef, :272) is +c
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, is_max, trace_cap=100000)
    st, res, trace = _ffi.solve(tab, basis, is_max,
                                _ffi.make_opts(writeback_full=True, trace_capacity=100000))
    assert st == ost and res.iterations == oit and trace == otrace
    assert np.array_equal(tab, o_tab) and np.array_equal(basis, o_basis)
    assert res.objective == o_tab[-1, -1]


def test_config2_m1024_n2048_objective_and_basis():
    """BASELINE config 2: objective within 1e-8 of the reference rule's answer (here: bit-exact)."""
    tab, basis = synthetic.dense_tableau(1024, 2048, seed=1234)
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, parallel=True, trace_cap=1 << 16)
    st, res, trace = _ffi.solve(tab, basis, True, _ffi.make_opts(writeback_full=True,
                                                                 trace_capacity=1 << 16))
    assert st == ost == _ffi.OK and res.iterations == oit
    assert trace == otrace
    assert np.array_equal(basis, o_basis)
    assert np.array_equal(tab, o_tab)
    assert abs(res.objective - 545.8113461511593) <= 1e-8 * 545.8113461511593   # HiGHS, SURVEY 6


def test_padded_host_leading_dimension(small_path):
    tab0, basis = random_tableau(50, 70, seed=5)
    wide = np.full((tab0.shape[0], tab0.shape[1] + 13), np.nan)
    wide[:, :tab0.shape[1]] = tab0
    view = wide[:, :tab0.shape[1]]
    o_tab, o_basis = tab0.copy(), basis.copy()
    oracle.solve(o_tab, o_basis, True)
    st, res, _ = _ffi.solve(view, basis, True, _ffi.make_opts(writeback_full=True))
    assert st == _ffi.OK and np.array_equal(view, o_tab) and np.array_equal(basis, o_basis)
    assert np.isnan(wide[:, tab0.shape[1]:]).all()


def test_iteration_limit_and_resume():
    tab, basis = random_tableau(120, 200, seed=9)
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, trace_cap=4096)
    assert oit > 20
    with _ffi.DeviceTableau(*tab.shape, opts=_ffi.make_opts(trace_capacity=4096)) as d:
        d.upload(tab, basis)
        st, res, _ = d.iterate(7)
        assert st == _ffi.ITERATION_LIMIT and res.iterations == 7
        st, res, _ = d.iterate(5)
        assert st == _ffi.ITERATION_LIMIT and res.iterations == 5
        st, res, trace = d.iterate(0)
        assert st == _ffi.OK and res.iterations == oit - 12 and trace == otrace
        g_tab, g_basis = d.download()
    assert np.array_equal(g_tab, o_tab) and np.array_equal(g_basis, o_basis)
    # one-shot call with opts.max_iters
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(max_iters=3))
    assert st == _ffi.ITERATION_LIMIT and res.iterations == 3


def test_fp_tolerance_keyword_changes_thresholds_like_the_oracle(small_path):
    tab, basis = random_tableau(60, 90, seed=21)
    for tol in (1.0, 1024.0, 1e9):
        o_tab, o_basis = tab.copy(), basis.copy()
        ost, oit, otrace = oracle.solve(o_tab, o_basis, True, tol=tol, trace_cap=4096)
        g_tab, g_basis = tab.copy(), basis.copy()
        st, res, trace = _ffi.solve(g_tab, g_basis, True,
                                    _ffi.make_opts(fp_tolerance=tol, writeback_full=True,
                                                   trace_capacity=4096))
        assert (st, res.iterations, trace) == (ost, oit, otrace)
        assert np.array_equal(g_tab, o_tab)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 20, 21, 22, 23, 30])
def test_every_pivot_kernel_variant_is_bit_exact(variant):
    """1..13: one launch per pivot, every tile shape; 20..23: the persistent cooperative loop
    (k_persist), every tile-role variant; 30: the tile staged through shared memory by the bulk
    asynchronous copy engine (k_iter2_bulk: cp.async.bulk + mbarrier)."""
    tab, basis = random_tableau(150, 333, seed=77)     # odd C, rows not a multiple of any tile
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, _ = oracle.solve(o_tab, o_basis, True)
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(writeback_full=True,
                                                             pivot_variant=variant))
    assert st == ost and res.iterations == oit and np.array_equal(tab, o_tab)
    assert res.loop_mode == (2 if 20 <= variant < 30 else 1)


@pytest.mark.parametrize("look", ["1", "2"], ids=["round-1-look-role", "two-phase-look-role"])
@pytest.mark.parametrize("look_ctas", [1, 3, 16])
@pytest.mark.parametrize("m,n,rule", [(700, 3500, 0), (90, 5000, 1), (3000, 200, 0)])
def test_per_pivot_loop_both_look_roles_bit_exact(m, n, rule, look_ctas, look, monkeypatch):
    """One launch per pivot (variant 10) with either look role: k_iter (round 1: three stages, the
    objective row re-read from the tableau) and k_iter2 (persist.cuh's two phases on compact
    copies, one decision per launch, state carried between launches)."""
    monkeypatch.setenv("B200LP_LOOK", look)
    monkeypatch.setenv("B200LP_LOOK_CTAS", str(look_ctas))
    tab, basis = random_tableau(m, n, seed=m + n, signed=True)
    o_tab, o_basis = tab.copy(), basis.copy()
    cap = 600
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=cap, trace_cap=cap,
                                    parallel=True)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(pivot_rule=rule, max_iters=cap, trace_capacity=cap,
                                               writeback_full=True, pivot_variant=10))
    assert (st, res.iterations) == (ost, oit) and trace == otrace
    assert res.loop_mode == 1
    assert np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


@pytest.mark.parametrize("cluster", ["1", "0"], ids=["cluster-barrier", "global-barrier"])
@pytest.mark.parametrize("look_ctas", [2, 4, 8])
@pytest.mark.parametrize("m,n,rule", [(700, 3500, 0), (90, 5000, 1), (3000, 200, 0)])
def test_persistent_loop_look_grid_as_a_cluster_is_bit_exact(m, n, rule, look_ctas, cluster, monkeypatch):
    """k_persist with its look grid launched as one thread-block cluster (partial argmins through
    distributed shared memory, barrier.cluster) and with the global-memory barriers: same bits."""
    monkeypatch.setenv("B200LP_LOOK_CTAS", str(look_ctas))
    monkeypatch.setenv("B200LP_CLUSTER", cluster)
    tab, basis = random_tableau(m, n, seed=m + n, signed=True)
    o_tab, o_basis = tab.copy(), basis.copy()
    cap = 600
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=cap, trace_cap=cap,
                                    parallel=True)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(pivot_rule=rule, max_iters=cap, trace_capacity=cap,
                                               writeback_full=True, pivot_variant=20))
    assert (st, res.iterations) == (ost, oit) and trace == otrace
    assert res.loop_mode == 2 and res.look_ctas == look_ctas
    assert res.look_cluster == (look_ctas if cluster == "1" else 0)
    assert np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


@pytest.mark.parametrize("look_ctas", [1, 2, 5, 16])
@pytest.mark.parametrize("m,n,rule", [(700, 3500, 0), (90, 5000, 1), (3000, 200, 0)])
def test_persistent_loop_any_look_grid_is_bit_exact(m, n, rule, look_ctas, monkeypatch):
    """k_persist with 1..16 look CTAs (the look-grid barriers and the per-CTA partial argmins):
    status, pivot trace, basis and every cell equal the oracle's."""
    monkeypatch.setenv("B200LP_LOOK_CTAS", str(look_ctas))
    tab, basis = random_tableau(m, n, seed=m + n, signed=True)
    o_tab, o_basis = tab.copy(), basis.copy()
    cap = 600
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=cap, trace_cap=cap,
                                    parallel=True)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(pivot_rule=rule, max_iters=cap, trace_capacity=cap,
                                               writeback_full=True, pivot_variant=20))
    assert (st, res.iterations) == (ost, oit) and trace == otrace
    assert res.loop_mode == 2 and res.look_ctas == look_ctas
    assert np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


# ------------------------------------------------------------------ degenerate / Bland (config 5)
def test_beale_cycles_under_reference_rule_and_bland_terminates(small_path):
    g = G.BEALE
    tab, basis = f64(g["matrix"]), i32(g["basis"])
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=0, max_iters=300, trace_cap=300)
    st, res, trace = _ffi.solve(tab.copy(), basis.copy(), True,
                                _ffi.make_opts(max_iters=300, trace_capacity=300))
    assert st == ost and res.iterations == oit and trace == otrace
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=1, max_iters=300, trace_cap=300)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(pivot_rule=_ffi.RULE_BLAND, max_iters=300,
                                               trace_capacity=300, writeback_full=True))
    assert st == ost == _ffi.OK and trace == otrace and np.array_equal(tab, o_tab)
    assert abs(res.objective - 0.05) < 1e-12


@pytest.mark.parametrize("rule", [0, 1])
def test_degenerate_lp_basis_bit_exact(rule):
    """Config 5 at reduced size: exact ratio ties (zero-RHS cone rows, small-integer data)."""
    tab, basis = synthetic.dense_tableau(128, 128, seed=1234, degenerate=True)
    o_tab, o_basis = tab.copy(), basis.copy()
    cap = 20000
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=cap, trace_cap=cap)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(pivot_rule=rule, max_iters=cap, trace_capacity=cap,
                                               writeback_full=True))
    assert st == ost and res.iterations == oit
    assert trace == otrace and np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


# ------------------------------------------------------------------ two-phase, randomized
@pytest.mark.parametrize("feas_mode", [0, 1], ids=["scaled", "reference"])
@pytest.mark.parametrize("seed", range(40))
def test_two_phase_ratio_problems_match_oracle_in_both_feasibility_modes(seed, feas_mode, small_path):
    """The transition's "is it zero" tests (src/simplex.lisp:405-434) in both modes
    (include/b200lp.h B200LP_FEAS_*): status, pivot counts, redundant rows and both tableaus equal
    the oracle's, bit for bit -- including the LPs the literal mode rejects."""
    import random_problems
    from linear_programming_b200 import problem as P, simplex
    objective, forms, names, ref = random_problems.generate_ratio(seed)
    problem = P.make_linear_problem(objective, *forms)
    built = simplex.build_tableau(problem, problem)
    if not isinstance(built, list):
        pytest.skip("no artificial rows")
    art, main = built
    o = [x.copy() for x in (art.matrix, art.basis_columns, main.matrix, main.basis_columns)]
    ost, oits = oracle.solve_two_phase(*o, False, feas_mode=feas_mode, with_redundant=True)
    st, res = _ffi.solve_two_phase(art.matrix, art.basis_columns, main.matrix, main.basis_columns,
                                   False, _ffi.make_opts(writeback_full=True, feas_mode=feas_mode))
    assert st == ost
    if st in (_ffi.OK, _ffi.UNBOUNDED):
        assert (res.iterations_phase1, res.iterations_cleanup, res.iterations, res.redundant_rows) == oits
        assert np.array_equal(main.matrix, o[2]) and np.array_equal(main.basis_columns, o[3])
        assert np.array_equal(art.matrix, o[0]) and np.array_equal(art.basis_columns, o[1])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_two_phase_random_matches_oracle(seed, small_path):
    """>= and = rows need artificials (src/simplex.lisp:258-263, 288-325)."""
    rng = np.random.default_rng(seed)
    m, n = 40, 30
    A = rng.integers(1, 9, size=(m, n)).astype(np.float64)
    x0 = rng.integers(0, 4, size=n).astype(np.float64)
    kinds = rng.integers(0, 3, size=m)            # 0: <=, 1: >=, 2: =
    rhs = A @ x0 + np.where(kinds == 0, 5.0, np.where(kinds == 1, -3.0, 0.0))
    rhs = np.maximum(rhs, 0.0)
    c = rng.integers(1, 9, size=n).astype(np.float64)
    n_slack = int((kinds != 2).sum())
    art_rows = [i for i in range(m) if kinds[i] != 0]
    C = n + n_slack + 1
    main = np.zeros((m + 1, C))
    mb = np.zeros(m, np.int32)
    off = 0
    for i in range(m):
        main[i, :n] = A[i]
        main[i, -1] = rhs[i]
        if kinds[i] == 0:
            main[i, n + off] = 1.0; mb[i] = n + off; off += 1
        elif kinds[i] == 1:
            main[i, n + off] = -1.0; mb[i] = C; off += 1
        else:
            mb[i] = C
    main[m, :n] = -c
    na = len(art_rows)
    art = np.zeros((m + 1, C + na))
    ab = mb.copy()
    art[:m, :C - 1] = main[:m, :C - 1]
    art[:m, -1] = main[:m, -1]
    # reference pushes rows, so artificial columns are assigned in reverse row order (:258, 295-299)
    for k, row in enumerate(reversed(art_rows)):
        art[row, C - 1 + k] = 1.0
        ab[row] = C - 1 + k
    art[m, :C - 1] = art[art_rows][:, :C - 1].sum(axis=0)
    art[m, -1] = art[art_rows][:, -1].sum()
    o = [x.copy() for x in (art, ab, main, mb)]
    ost, oits = oracle.solve_two_phase(*o, True)
    st, res = _ffi.solve_two_phase(art, ab, main, mb, True, _ffi.make_opts(writeback_full=True))
    assert st == ost
    assert (res.iterations_phase1, res.iterations_cleanup, res.iterations) == oits
    if st == _ffi.OK:
        assert np.array_equal(main, o[2]) and np.array_equal(mb, o[3])
        assert np.array_equal(art, o[0]) and np.array_equal(ab, o[1])


# ------------------------------------------------------------------ argument errors stay on this side
def test_invalid_arguments_raise_not_crash():
    tab, basis = random_tableau(8, 8, seed=1)
    with _ffi.DeviceTableau(*tab.shape) as d:
        d.upload(tab, basis)
        with pytest.raises(_ffi.B200DeviceError):
            d.pivot(0, 99)
        with pytest.raises(_ffi.B200DeviceError):
            d.find_pivoting_row(10 ** 6)
    with pytest.raises(_ffi.B200DeviceError):
        _ffi.DeviceTableau(1, 1)


# ------------------------------------------------------------------ BASELINE.json's full sizes
def _basis_columns_are_unit_vectors(tab, basis):
    """Size-independent invariant of n-pivot-row (src/simplex.lisp:344-357): after any number of
    pivots the basic column of row i is EXACTLY e_i (x/x = 1, a - a*1 = 0), objective row included."""
    m = tab.shape[0] - 1
    cols = tab[:, basis]                                   # (m+1) x m
    want = np.zeros_like(cols)
    want[np.arange(m), np.arange(m)] = 1.0
    return np.array_equal(cols, want)


@pytest.mark.parametrize("m,n,degenerate,rule,k_exact", [
    (8192, 16384, False, 0, 10),       # config 3
    (4096, 4096, True, 0, 150),        # config 5, reference rule
    (4096, 4096, True, 1, 150),        # config 5, Bland
    (16384, 32768, False, 0, 50),      # config 4 on one GPU
])
def test_full_size_configs_prefix_bit_exact_then_invariants(m, n, degenerate, rule, k_exact):
    tab, basis = synthetic.dense_tableau(m, n, seed=1234, degenerate=degenerate)
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=k_exact,
                                    parallel=True, trace_cap=k_exact)
    opts = _ffi.make_opts(pivot_rule=rule, trace_capacity=4096)
    with _ffi.DeviceTableau(*tab.shape, opts=opts) as d:
        d.upload(tab, basis)
        st, res, trace = d.iterate(k_exact)
        assert (st, res.iterations) == (ost, oit) and trace == otrace
        rhs, obj, g_basis = d.download_solution()
        assert np.array_equal(rhs, o_tab[:, -1]) and np.array_equal(obj, o_tab[-1])
        assert np.array_equal(g_basis, o_basis)
        del o_tab
        # keep going well past what the CPU can check cell by cell
        before = res.objective
        st, res, trace = d.iterate(400)
        assert st in (_ffi.ITERATION_LIMIT, _ffi.OK)
        assert res.objective >= before                      # max problem: never decreases
        full, g_basis = d.download()
    assert len(set(g_basis.tolist())) == m                  # a basis: m distinct columns
    assert _basis_columns_are_unit_vectors(full, g_basis)
    assert full[-1, -1] == res.objective
    if not degenerate:
        assert (full[:-1, -1] >= -1e-9).all()               # primal feasibility is kept


@pytest.mark.parametrize("m,n,degenerate", [(1024, 2048, False), (8192, 16384, False),
                                            (4096, 4096, True)])
def test_full_solve_is_certified_optimal_at_baseline_sizes(m, n, degenerate):
    """Configs 2, 3 and 5 (moderately degenerate variant) solved to optimality through
    b200lp_solve, then certified by LP duality on the CPU."""
    A, b, c = synthetic.dense_lp(m, n, seed=1234, degenerate=degenerate, zero_frac=1 / 32)
    tab, basis = synthetic.tableau_from_lp(A, b, c)
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(max_iters=400000))
    assert st == _ffi.OK
    value = certify_optimal(A, b, c, tab[:, -1], tab[-1], basis)
    assert value == res.objective
    if (m, n) == (1024, 2048):
        assert abs(value - 545.8113461511593) <= 1e-8 * 545.8113461511593      # HiGHS, SURVEY 6


@pytest.mark.parametrize("fixture,m,n", [("cfg2_final", 1024, 2048), ("cfg3_final", 8192, 16384),
                                         ("cfg5_final", 4096, 4096), ("cfg5_rule1_final", 4096, 4096)])
def test_full_solve_end_state_equals_the_oracle_fixture(fixture, m, n):
    """BASELINE configs 2, 3 and 5 through b200lp_solve to where the ORACLE ends: pivot count, the
    whole pivot trace, final basis, RHS column and objective row are equal bit for bit.
    tests/golden/<fixture>.npz is generated by tools/make_full_goldens.py (18 751 pivots of config 3
    take the CPU a quarter of an hour, so the end state travels as a fixture).
      cfg2 / cfg3        the reference rule to optimality (696 / 18 751 pivots)
      cfg5               m = n = 4096 degenerate (64 cone rows through the origin, small-integer data,
                         exact ratio ties), the reference rule to optimality: 53 616 pivots
      cfg5, Bland        the same LP under Bland's rule, compared at a 50 000-pivot cap -- Bland
                         needs on the order of 10^7 pivots here, far beyond what
                         the CPU oracle can follow"""
    import os
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{fixture}.npz"))
    assert (int(fx["m"]), int(fx["n"]), int(fx["seed"])) == (m, n, 1234)
    rule, status, iters = int(fx["rule"]), int(fx["status"]), int(fx["iterations"])
    tab, basis = synthetic.dense_tableau(m, n, seed=1234, degenerate=bool(fx["degenerate"]),
                                         zero_frac=float(fx["zero_frac"]))
    capped = status == _ffi.ITERATION_LIMIT
    st, res, trace = _ffi.solve(tab, basis, True, _ffi.make_opts(
        trace_capacity=iters + 16, pivot_rule=rule, max_iters=iters if capped else 0))
    assert st == status and res.iterations == iters
    assert np.array_equal(np.asarray(trace, np.int32).reshape(-1, 2), fx["trace"])
    assert np.array_equal(basis, fx["basis"])
    assert np.array_equal(tab[:, -1], fx["rhs"]) and np.array_equal(tab[-1], fx["obj_row"])
    assert res.objective == float(fx["objective"])


@pytest.mark.parametrize("m,rule", [(1024, 0), (512, 1)])
def test_degenerate_full_solve_final_basis_bit_exact(m, rule):
    """Thousands of exact ratio ties, solved to the end under both rules: the final basis, the
    whole tableau and the pivot count equal the oracle's."""
    tab, basis = synthetic.dense_tableau(m, m, seed=1234, degenerate=True, zero_frac=1 / 32)
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, _ = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=400000, parallel=True)
    st, res, _ = _ffi.solve(tab, basis, True, _ffi.make_opts(pivot_rule=rule, max_iters=400000,
                                                             writeback_full=True))
    assert st == ost == _ffi.OK and res.iterations == oit
    assert np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


# ------------------------------------------------------------------ randomized shape sweep
def _sweep_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for k in range(40):                                   # tiny to small, every alignment of C
        cases.append((int(rng.integers(1, 70)), int(rng.integers(1, 90)), k))
    cases += [(5, 10, 100), (15, 16, 101), (31, 32, 102)]            # C = n+m+1 multiple of 16: no pad
    cases += [(3, 5000, 103), (1, 9000, 104), (2, 4200, 105),       # long rows, fewer rows than look CTAs
              (700, 3500, 106), (1500, 40, 107), (4100, 3, 108)]    # multi-CTA look, tall and thin
    return cases


@pytest.mark.parametrize("small", ["1", "0"], ids=["one-cta-path-on", "one-cta-path-off"])
@pytest.mark.parametrize("m,n,seed", _sweep_cases())
def test_random_shapes_signs_rules_bit_exact(m, n, seed, small, monkeypatch):
    """Mixed-sign data (optimal, unbounded and stalled outcomes all occur), max and min problems,
    both pivot rules, every row-stride alignment; status, pivot trace, basis and every cell must
    equal the oracle's."""
    monkeypatch.setenv("B200LP_SMALL", small)       # "0": tableaus that fit one CTA take the big path too
    rng = np.random.default_rng(seed)
    A = rng.random((m, n)) - (0.0 if seed % 3 else 0.3)
    if seed % 4 == 0:
        A = np.round(A * 4)                               # small integers: exact ties
    b = rng.uniform(n / 8.0, 3.0 * n / 8.0, m) * (rng.random(m) > (0.2 if seed % 5 == 0 else -1))
    c = rng.random(n) - (0.0 if seed % 2 else 0.2)
    tab, basis = synthetic.tableau_from_lp(A, b, c)
    is_max = seed % 7 != 0
    rule = 1 if seed % 6 == 0 else 0
    if not is_max:
        tab[-1, :n] *= -1.0
    cap = 3000
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, is_max, rule=rule, max_iters=cap, trace_cap=cap,
                                    parallel=m * n > 100000)
    st, res, trace = _ffi.solve(tab, basis, is_max,
                                _ffi.make_opts(pivot_rule=rule, max_iters=cap, trace_capacity=cap,
                                               writeback_full=True))
    assert (st, res.iterations) == (ost, oit)
    assert trace == otrace
    assert np.array_equal(basis, o_basis) and np.array_equal(tab, o_tab)


def test_pooled_handles_are_rebound_cleanly_across_shapes(small_path):
    """The one-shot calls reuse idle handles for nearby shapes; nothing of a previous solve (ring
    slots, ping-pong parity, basis, trace, tolerances, rule) may leak into the next."""
    shapes = [(30, 40), (31, 41), (12, 90), (33, 38), (2, 3), (60, 100), (30, 40), (64, 190), (5, 7)]
    for k, (m, n) in enumerate(shapes * 2):
        tab, basis = random_tableau(m, n, seed=500 + k, signed=k % 3 == 0)
        rule, tol, cap = k % 2, (1024.0, 64.0)[k % 2], (0, 512, 4096)[k % 3]
        o_tab, o_basis = tab.copy(), basis.copy()
        ost, oit, otrace = oracle.solve(o_tab, o_basis, True, tol=tol, rule=rule, max_iters=2000,
                                        trace_cap=cap)
        st, res, trace = _ffi.solve(tab, basis, True,
                                    _ffi.make_opts(fp_tolerance=tol, pivot_rule=rule, max_iters=2000,
                                                   trace_capacity=cap, writeback_full=True))
        assert (st, res.iterations) == (ost, oit) and trace == otrace[:cap]
        assert np.array_equal(tab, o_tab) and np.array_equal(basis, o_basis)
        if k == 9:
            _ffi.shutdown()                                # dropping the pool is always safe
    g = G.EQ_SOLVED                                        # two-phase draws two handles from the pool
    b = g["initial"]
    for _ in range(3):
        art, ab = f64(b["art_matrix"]), i32(b["art_basis"])
        main, mb = f64(b["main_matrix"]), i32(b["main_basis"])
        st, res = _ffi.solve_two_phase(art, ab, main, mb, True, _ffi.make_opts(writeback_full=True))
        assert st == _ffi.OK and mb.tolist() == g["main_basis"] and res.objective == 28.5
