"""Pins the oracle (exact Fraction tier and fp64 C tier) to the reference's golden tableaus.

Mirrors t/simplex.lisp's tests `pivot-row`, `basic-problem`, `equality-constraint`,
`leq-constraint`, `unsolvable-problems` and t/integration.lisp `basic-problem`.
CPU only (-m "not gpu").
"""
import copy
from fractions import Fraction

import numpy as np
import pytest

from golden import reference_goldens as G
from oracle import exact, oracle


def f64(rows):
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float64)


def i32(b):
    return np.array(b, dtype=np.int32)


# ----------------------------------------------------------------------------- exact tier
def test_exact_single_pivot():
    g = G.SINGLE_PIVOT
    tab, basis = copy.deepcopy(g["initial"]["matrix"]), list(g["initial"]["basis"])
    exact.n_pivot_row(tab, basis, g["col"], g["row"])
    assert tab == g["matrix"] and basis == g["basis"] and tab[-1][-1] == g["objective"]


def test_exact_basic_solve():
    g = G.BASIC_SOLVED
    tab, basis = copy.deepcopy(g["initial"]["matrix"]), list(g["initial"]["basis"])
    trace = []
    st, it = exact.n_solve_tableau(tab, basis, True, trace=trace)
    assert (st, it) == (exact.OPTIMAL, g["pivots"])
    assert tab == g["matrix"] and basis == g["basis"] and trace == g["trace"]
    assert tab[-1][-1] == Fraction(57, 2)


@pytest.mark.parametrize("g", [G.EQ_SOLVED, G.GEQ_SOLVED], ids=["eq", "geq"])
def test_exact_two_phase(g):
    b = g["initial"]
    art, ab = copy.deepcopy(b["art_matrix"]), list(b["art_basis"])
    main, mb = copy.deepcopy(b["main_matrix"]), list(b["main_basis"])
    st, its = exact.n_solve_two_phase(art, ab, main, mb, True)
    assert st == exact.OPTIMAL and its == g["pivots"]
    assert art == g["art_matrix"] and ab == g["art_basis"]
    assert main == g["main_matrix"] and mb == g["main_basis"]
    assert main[-1][-1] == g["objective"]


def test_exact_infeasible_unbounded():
    g = G.INFEASIBLE
    st, _ = exact.n_solve_two_phase(copy.deepcopy(g["art_matrix"]), list(g["art_basis"]),
                                    copy.deepcopy(g["main_matrix"]), list(g["main_basis"]), True)
    assert st == exact.INFEASIBLE
    g = G.UNBOUNDED
    st, _ = exact.n_solve_tableau(copy.deepcopy(g["matrix"]), list(g["basis"]), True)
    assert st == exact.UNBOUNDED


def test_exact_assembly():
    g = G.ASSEMBLY
    tab, basis = copy.deepcopy(g["matrix"]), list(g["basis"])
    st, it = exact.n_solve_tableau(tab, basis, True)
    assert (st, it) == (exact.OPTIMAL, g["pivots"]) and tab[-1][-1] == g["objective"]
    names = ["widgets", "d1", "d2", "d3"]
    for col, name in enumerate(names):
        val = tab[basis.index(col)][-1] if col in basis else 0
        assert val == g["primal"][name]
        lo, hi = g["bounds"][name]
        assert lo <= float(val) <= hi


def test_exact_beale_cycles_under_reference_rule_and_bland_terminates():
    g = G.BEALE
    tab, basis = copy.deepcopy(g["matrix"]), list(g["basis"])
    st, it = exact.n_solve_tableau(tab, basis, True, rule=0, max_iters=600)
    assert st == exact.ITERATION_LIMIT          # Dantzig + lowest-index ties cycles
    tab, basis = copy.deepcopy(g["matrix"]), list(g["basis"])
    st, it = exact.n_solve_tableau(tab, basis, True, rule=1, max_iters=600)
    assert st == exact.OPTIMAL and tab[-1][-1] == g["objective"]


# -------------------------------------------------------------------------- fp64 C tier
def test_c_epsilon_is_cl_double_float_epsilon():
    assert oracle.cl_epsilon() == float.fromhex("0x1.0000000000001p-53")
    assert oracle.cl_epsilon() * 128 == 1.4210854715202007e-14
    assert oracle.cl_epsilon() * 512 == 5.684341886080803e-14


def test_c_single_pivot_bit_exact():
    g = G.SINGLE_PIVOT
    tab, basis = f64(g["initial"]["matrix"]), i32(g["initial"]["basis"])
    oracle.pivot(tab, basis, g["col"], g["row"])
    assert np.array_equal(tab, f64(g["matrix"])) and basis.tolist() == g["basis"]


def test_c_basic_solve_bit_exact():
    g = G.BASIC_SOLVED
    tab, basis = f64(g["initial"]["matrix"]), i32(g["initial"]["basis"])
    st, it, trace = oracle.solve(tab, basis, True, trace_cap=16)
    assert (st, it, trace) == (oracle.OPTIMAL, g["pivots"], g["trace"])
    assert np.array_equal(tab, f64(g["matrix"]))       # halves are exact in fp64
    assert basis.tolist() == g["basis"] and tab[-1, -1] == 28.5


def test_c_padded_leading_dimension():
    g = G.BASIC_SOLVED
    wide = np.full((3, 16), 7.0)
    wide[:, :6] = f64(g["initial"]["matrix"])
    view = wide[:, :6]
    basis = i32(g["initial"]["basis"])
    lib = oracle.lib()
    import ctypes
    it = ctypes.c_int64()
    st = lib.oracle_solve(wide.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 3, 6, 16,
                          basis.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1, 1024.0, 0, 0, 0,
                          ctypes.byref(it), None, None, 0)
    assert st == 0 and it.value == 2
    assert np.array_equal(view, f64(g["matrix"])) and np.all(wide[:, 6:] == 7.0)


@pytest.mark.parametrize("g", [G.EQ_SOLVED, G.GEQ_SOLVED], ids=["eq", "geq"])
def test_c_two_phase(g):
    b = g["initial"]
    art, ab = f64(b["art_matrix"]), i32(b["art_basis"])
    main, mb = f64(b["main_matrix"]), i32(b["main_basis"])
    st, its = oracle.solve_two_phase(art, ab, main, mb, True)
    assert st == oracle.OPTIMAL and its == g["pivots"]
    assert ab.tolist() == g["art_basis"] and mb.tolist() == g["main_basis"]
    np.testing.assert_allclose(art, f64(g["art_matrix"]), rtol=0, atol=1e-14)
    np.testing.assert_allclose(main, f64(g["main_matrix"]), rtol=1e-14, atol=1e-14)
    assert abs(main[-1, -1] - float(g["objective"])) <= 1e-8 * abs(float(g["objective"]))


def test_c_infeasible_unbounded():
    g = G.INFEASIBLE
    st, _ = oracle.solve_two_phase(f64(g["art_matrix"]), i32(g["art_basis"]),
                                   f64(g["main_matrix"]), i32(g["main_basis"]), True)
    assert st == oracle.INFEASIBLE
    g = G.UNBOUNDED
    st, _, _ = oracle.solve(f64(g["matrix"]), i32(g["basis"]), True)
    assert st == oracle.UNBOUNDED


def test_c_assembly_within_reference_intervals():
    g = G.ASSEMBLY
    tab, basis = f64(g["matrix"]), i32(g["basis"])
    st, it, _ = oracle.solve(tab, basis, True)
    assert (st, it) == (oracle.OPTIMAL, g["pivots"])
    assert abs(tab[-1, -1] - float(g["objective"])) <= 1e-8 * float(g["objective"])
    lo, hi = g["bounds"]["revenue"]
    assert lo <= tab[-1, -1] <= hi


def test_c_matches_exact_on_random_small_lps():
    """Tier-1 vs tier-2: same pivot trace and objective within 1e-8 on small random LPs."""
    rng = np.random.default_rng(7)
    for trial in range(12):
        m, n = int(rng.integers(3, 14)), int(rng.integers(3, 14))
        A = rng.integers(0, 10, size=(m, n)).astype(np.float64)
        b = rng.integers(1, 50, size=m).astype(np.float64)
        c = rng.integers(1, 10, size=n).astype(np.float64)
        tab = np.zeros((m + 1, n + m + 1))
        tab[:m, :n], tab[:m, n:n + m], tab[:m, -1], tab[m, :n] = A, np.eye(m), b, -c
        basis = np.arange(n, n + m, dtype=np.int32)
        ex = exact.to_fraction_matrix(tab.tolist())
        eb = basis.tolist()
        et = []
        est, eit = exact.n_solve_tableau(ex, eb, True, trace=et, max_iters=500)
        st, it, trace = oracle.solve(tab, basis, True, trace_cap=500, max_iters=500)
        if est != st or et != trace:
            # an exact tie broken by fp64 round-off is legitimate; objective must still agree
            assert st == est == oracle.OPTIMAL
        if st == oracle.OPTIMAL:
            ref = float(ex[-1][-1])
            assert abs(tab[-1, -1] - ref) <= 1e-8 * max(1.0, abs(ref))


def test_c_parallel_rows_bit_identical():
    rng = np.random.default_rng(3)
    m, n = 96, 160
    tab = np.zeros((m + 1, n + m + 1))
    tab[:m, :n] = rng.random((m, n))
    tab[:m, n:n + m] = np.eye(m)
    tab[:m, -1] = rng.uniform(n / 8, 3 * n / 8, m)
    tab[m, :n] = -rng.random(n)
    t1, b1 = tab.copy(), np.arange(n, n + m, dtype=np.int32)
    t2, b2 = tab.copy(), b1.copy()
    r1 = oracle.solve(t1, b1, True, parallel=False, trace_cap=4096)
    r2 = oracle.solve(t2, b2, True, parallel=True, trace_cap=4096)
    assert r1 == r2 and np.array_equal(t1, t2) and np.array_equal(b1, b2)


def test_c_bland_on_beale():
    g = G.BEALE
    tab, basis = f64(g["matrix"]), i32(g["basis"])
    st, it, _ = oracle.solve(tab, basis, True, rule=1, max_iters=600)
    assert st == oracle.OPTIMAL and abs(tab[-1, -1] - 0.05) < 1e-12


@pytest.mark.parametrize("m,n,degenerate", [(200, 300, False), (256, 256, True)])
def test_c_oracle_optimum_is_certified_by_duality(m, n, degenerate):
    """Independent of any golden: primal + dual feasibility and a closed duality gap."""
    from lp_checks import certify_optimal
    from linear_programming_b200 import synthetic
    A, b, c = synthetic.dense_lp(m, n, seed=3, degenerate=degenerate, zero_frac=1 / 32)
    tab, basis = synthetic.tableau_from_lp(A, b, c)
    st, _, _ = oracle.solve(tab, basis, True, max_iters=200000)
    assert st == oracle.OPTIMAL
    certify_optimal(A, b, c, tab[:, -1], tab[-1], basis)


def test_cycle_probe_proves_beale_cycles_under_the_reference_rule_and_not_under_bland():
    """oracle.solve_cycle_probe: a revisited basis under a deterministic rule is a proven cycle.
    Beale's example cycles with period 6 under the reference's rule (which has no anti-cycling,
    src/simplex.lisp:455-460); Bland's rule terminates without ever revisiting a basis."""
    import numpy as np
    from golden import reference_goldens as G
    from oracle import oracle
    g = G.BEALE
    f64 = lambda rows: np.array([[float(x) for x in r] for r in rows], dtype=np.float64)   # noqa: E731
    tab, basis = f64(g["matrix"]), np.array(g["basis"], np.int32)
    st, info = oracle.solve_cycle_probe(tab.copy(), basis.copy(), True, rule=0, max_iters=300)
    assert st == oracle.ITERATION_LIMIT and info["revisit_at"] == 6 and info["first_visit_at"] == 0
    assert info["degenerate_pivots"] == info["pivots"] == 300
    st, info = oracle.solve_cycle_probe(tab.copy(), basis.copy(), True, rule=1, max_iters=300)
    assert st == oracle.OPTIMAL and info["revisit_at"] == -1


def test_cycle_probe_on_the_degenerate_generator_stalls_but_never_cycles():
    """Config 5's family at a size where it runs out: thousands of degenerate pivots, no basis
    ever revisited (so the reference rule stalls on it, it does not cycle), both rules reach the
    same optimum."""
    from linear_programming_b200 import synthetic
    from oracle import oracle
    objs = []
    for rule in (0, 1):
        tab, basis = synthetic.dense_tableau(256, 256, degenerate=True, zero_frac=0.5)
        st, info = oracle.solve_cycle_probe(tab, basis, True, rule=rule, max_iters=200000, parallel=True)
        assert st == oracle.OPTIMAL and info["revisit_at"] == -1 and info["degenerate_pivots"] > 1000
        objs.append(tab[-1, -1])
    assert abs(objs[0] - objs[1]) <= 1e-6 * abs(objs[0])
