"""Row-block sharding on real GPUs (needs >= 2 B200s; skipped otherwise): the pivot sequence and
every tableau cell must be identical to the 1-GPU / oracle run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from linear_programming_b200 import _ffi, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        return _ffi.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("exchange,mode,loop", [("p2p", 2, "persist"), ("p2p", 2, "iter"), ("nccl", 1, "iter")])
@pytest.mark.parametrize("ndev,m,n", [(2, 200, 300), (2, 33, 20), (4, 130, 257), (8, 64, 96), (2, 300, 5000)])
def test_in_process_multi_device_bit_exact(ndev, m, n, exchange, mode, loop, monkeypatch):
    """Every way the candidate pivot rows travel: peer-mapped buffers written by the persistent
    cooperative loop (default; one speculative push + one flag wait per pivot), peer-mapped
    buffers written inside the per-pivot iteration kernel, and the NCCL all-gather fallback."""
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    monkeypatch.setenv("B200LP_EXCHANGE", exchange)
    monkeypatch.setenv("B200LP_LOOP", loop)
    monkeypatch.setenv("B200LP_PEER_TIMEOUT_MS", "5000")
    tab, basis = synthetic.dense_tableau(m, n, seed=13)
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, trace_cap=1 << 16)
    st, res, trace = _ffi.solve(tab, basis, True,
                                _ffi.make_opts(devices=list(range(ndev)), writeback_full=True,
                                               trace_capacity=1 << 16))
    assert st == ost and res.iterations == oit and res.n_devices == ndev
    assert res.exchange_mode == mode and res.loop_mode == (2 if loop == "persist" else 1)
    assert trace == otrace
    assert np.array_equal(tab, o_tab) and np.array_equal(basis, o_basis)


def _torchrun(nproc, *args, exchange="p2p", loop="persist"):
    env = dict(os.environ, B200LP_EXCHANGE=exchange, B200LP_PEER_TIMEOUT_MS="5000", B200LP_LOOP=loop)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py"), *map(str, args)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


@needs2
@pytest.mark.parametrize("exchange,loop", [("p2p", "persist"), ("p2p", "iter"), ("nccl", "iter")])
@pytest.mark.parametrize("args", [(300, 500), (128, 128, 1, "degenerate"), (128, 128, 0, "degenerate")])
def test_one_process_per_gpu(args, exchange, loop):
    out = _torchrun(2, *args, exchange=exchange, loop=loop)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.skipif(_ngpu() < 8, reason="needs 8 GPUs")
def test_eight_ranks():
    out = _torchrun(8, 1000, 1500)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def _two_phase_lp(seed, m, n):
    """A random LP with <=, >= and = rows as the (art, main) tableaus build-tableau produces
    (src/simplex.lisp:258-263, 288-325)."""
    rng = np.random.default_rng(seed)
    A = rng.integers(1, 9, size=(m, n)).astype(np.float64)
    x0 = rng.integers(0, 4, size=n).astype(np.float64)
    kinds = rng.integers(0, 3, size=m)            # 0: <=, 1: >=, 2: =
    rhs = np.maximum(A @ x0 + np.where(kinds == 0, 5.0, np.where(kinds == 1, -3.0, 0.0)), 0.0)
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
    for k, row in enumerate(reversed(art_rows)):
        art[row, C - 1 + k] = 1.0
        ab[row] = C - 1 + k
    art[m, :C - 1] = art[art_rows][:, :C - 1].sum(axis=0)
    art[m, -1] = art[art_rows][:, -1].sum()
    return art, ab, main, mb


@needs2
@pytest.mark.parametrize("feas_mode", [0, 1], ids=["scaled", "reference"])
@pytest.mark.parametrize("ndev,m,n,seed", [(2, 40, 30, 1), (2, 90, 60, 2), (4, 61, 45, 3), (8, 120, 80, 4),
                                           (2, 300, 200, 5)])
def test_sharded_two_phase_is_bit_identical_to_the_oracle(ndev, m, n, seed, feas_mode):
    """b200lp_solve_two_phase on row-block shards (src/simplex.lisp:402-452): phase 1, the
    clean-up pivots on the owning shard, the per-shard coefficient copy and the objective re-pricing
    handed from shard to shard in rank order give the oracle's tableaus bit for bit."""
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    art, ab, main, mb = _two_phase_lp(seed, m, n)
    o = [x.copy() for x in (art, ab, main, mb)]
    ost, oits = oracle.solve_two_phase(*o, True, feas_mode=feas_mode, with_redundant=True)
    st, res = _ffi.solve_two_phase(art, ab, main, mb, True,
                                   _ffi.make_opts(devices=list(range(ndev)), writeback_full=True,
                                                  feas_mode=feas_mode))
    assert st == ost and res.n_devices == ndev
    if st in (_ffi.OK, _ffi.UNBOUNDED):
        assert (res.iterations_phase1, res.iterations_cleanup, res.iterations, res.redundant_rows) == oits
        assert np.array_equal(main, o[2]) and np.array_equal(mb, o[3])
        assert np.array_equal(art, o[0]) and np.array_equal(ab, o[1])


@needs2
def test_two_phase_golden_on_two_devices():
    """t/simplex.lisp:196-237 through the sharded two-phase call."""
    from golden import reference_goldens as G
    g = G.EQ_SOLVED
    b = g["initial"]
    f64 = lambda rows: np.array([[float(x) for x in r] for r in rows])     # noqa: E731
    art, ab = f64(b["art_matrix"]), np.array(b["art_basis"], np.int32)
    main, mb = f64(b["main_matrix"]), np.array(b["main_basis"], np.int32)
    st, res = _ffi.solve_two_phase(art, ab, main, mb, True, _ffi.make_opts(devices=[0, 1]))
    assert st == _ffi.OK and res.objective == 28.5 and mb.tolist() == g["main_basis"]
    assert res.n_devices == 2


@needs2
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_step_by_step_api_on_a_multi_device_handle(exchange, monkeypatch):
    """find-entering-column / find-pivoting-row / n-pivot-row one call at a time on a handle that
    row-block shards over two GPUs of this process, against the oracle."""
    monkeypatch.setenv("B200LP_EXCHANGE", exchange)
    tab, basis = synthetic.dense_tableau(45, 70, seed=3)
    o_tab, o_basis = tab.copy(), basis.copy()
    with _ffi.DeviceTableau(*tab.shape, opts=_ffi.make_opts(devices=[0, 1])) as d:
        d.upload(tab, basis)
        for _ in range(12):
            j = d.find_entering_column()
            assert (j if j is not None else -1) == oracle.find_entering_column(o_tab, True)
            if j is None:
                break
            r = d.find_pivoting_row(j)
            assert (r if r is not None else -1) == oracle.find_pivoting_row(o_tab, o_basis, j)
            d.pivot(j, r)
            oracle.pivot(o_tab, o_basis, j, r)
        st, res, _ = d.iterate(0)                          # and the pipelined loop picks up from there
        oracle.solve(o_tab, o_basis, True)
        g_tab, g_basis = d.download()
    assert st == _ffi.OK and np.array_equal(g_tab, o_tab) and np.array_equal(g_basis, o_basis)
