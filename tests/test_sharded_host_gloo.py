"""World-size-2 (and 3) gloo tests of the N > 1 path on CPU: partition arithmetic, scatter /
assemble, unique-id broadcast and the candidate-row exchange protocol (emulated per rank with
oracle primitives, tests/shard_emulator.py).  The sharded pivot sequence must be identical to the
unsharded oracle's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from linear_programming_b200 import sharded, synthetic
    from oracle import oracle
    from shard_emulator import OracleShard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tab, basis = synthetic.dense_tableau(m, n, seed=5)
        ref_tab, ref_basis = tab.copy(), basis.copy()
        ost, oit, otrace = oracle.solve(ref_tab, ref_basis, True, trace_cap=1 << 16)
        traces = []

        def factory(*a, **k):
            d = OracleShard(*a, **k)
            traces.append(d.trace)
            return d

        st, it, rhs, obj, full_basis = sharded.solve_sharded(
            tab, basis, True, None, device_factory=factory, unique_id_factory=lambda: bytes(range(128)))
        ok = (st == ost and it == oit and traces[0] == otrace
              and np.array_equal(rhs, ref_tab[:, -1]) and np.array_equal(obj, ref_tab[-1])
              and np.array_equal(full_basis, ref_basis)
              and np.array_equal(tab[:, -1], ref_tab[:, -1]) and np.array_equal(basis, ref_basis))
        q.put((rank, bool(ok), it))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,m,n", [(2, 40, 60), (3, 31, 47), (2, 3, 5)])
def test_sharded_protocol_matches_unsharded_oracle(world, m, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    assert len({it for _, _, it in results}) == 1


def test_partition_matches_c_abi_and_covers_rows():
    sys.path.insert(0, ROOT)
    from linear_programming_b200 import _ffi, sharded
    for m in (1, 2, 7, 8, 1000, 8192, 16385):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for rank in range(world):
                b, e = sharded.partition(m, world, rank)
                assert (b, e) == _ffi.partition(m, world, rank)
                cover += list(range(b, e))
            assert cover == list(range(m))


def test_local_block_and_assemble_roundtrip():
    sys.path.insert(0, ROOT)
    from linear_programming_b200 import sharded, synthetic
    tab, basis = synthetic.dense_tableau(11, 6, seed=2)
    pieces = []
    for rank in range(4):
        blk, bb, (b, e) = sharded.local_block(tab, basis, 4, rank)
        assert blk.shape == (e - b + 1, tab.shape[1]) and np.array_equal(blk[-1], tab[-1])
        pieces.append((blk[:, -1].copy(), blk[-1].copy(), bb))
    rhs, obj, full = sharded.assemble_solution(pieces, 11)
    assert np.array_equal(rhs, tab[:, -1]) and np.array_equal(obj, tab[-1]) and np.array_equal(full, basis)
