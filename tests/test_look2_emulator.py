"""CPU check of the round-2 look role's algebra (tests/look2_emulator.py: compact objective-row /
RHS copies, decisions taken two tableaus back, speculative one-exchange sharding) against the
unsharded oracle -- pivot trace, final tableau and basis bit for bit.  No GPU involved."""
import numpy as np
import pytest

import look2_emulator
from linear_programming_b200 import synthetic
from oracle import oracle


def _lp(m, n, seed, signed):
    rng = np.random.default_rng(seed)
    A = rng.random((m, n)) - (0.25 if signed else 0.0)
    if seed % 3 == 0:
        A = np.round(A * 4)                                # small integers: exact ratio ties
    b = rng.uniform(n / 8.0, 3.0 * n / 8.0, m) * (rng.random(m) > 0.15)
    c = rng.random(n) - (0.2 if seed % 2 else 0.0)
    return synthetic.tableau_from_lp(A, b, c)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("m,n,seed,rule,is_max", [(24, 30, 1, 0, True), (40, 25, 2, 1, True),
                                                  (33, 50, 3, 0, True), (17, 60, 4, 0, False),
                                                  (64, 64, 6, 1, True)])
def test_look2_model_equals_the_oracle(m, n, seed, rule, is_max, world):
    tab, basis = _lp(m, n, seed, signed=seed % 2 == 0)
    if not is_max:
        tab[-1, :n] *= -1.0
    cap = 400
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, is_max, rule=rule, max_iters=cap, trace_cap=cap)
    st, trace, out, out_basis = look2_emulator.solve(tab, basis, is_max, rule=rule, max_iters=cap,
                                                     world=min(world, m))
    assert (st, len(trace)) == (ost, oit) and trace == otrace
    assert np.array_equal(out, o_tab) and np.array_equal(out_basis, o_basis)


def test_look2_model_on_the_degenerate_family_both_rules():
    for rule in (0, 1):
        tab, basis = synthetic.dense_tableau(48, 48, degenerate=True, zero_frac=0.25)
        o_tab, o_basis = tab.copy(), basis.copy()
        ost, oit, otrace = oracle.solve(o_tab, o_basis, True, rule=rule, max_iters=5000, trace_cap=5000)
        st, trace, out, out_basis = look2_emulator.solve(tab, basis, True, rule=rule, max_iters=5000, world=4)
        assert (st, trace) == (ost, otrace) and np.array_equal(out, o_tab) and np.array_equal(out_basis, o_basis)
