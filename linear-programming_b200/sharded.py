"""Row-block sharded solves, one process per GPU.

Host-side plumbing only: which rows a rank owns, how the tableau is scattered and how the
solution is put back together.  The data path (per-iteration candidate-row exchange and the
kernels) lives in libb200lp.so; `torch.distributed` is used here for the rendezvous (NCCL unique
id) and for gathering the few result vectors, never inside the iteration loop.

Partition rule = b200lp_partition: contiguous blocks of ceil(m / world) constraint rows; every
rank also holds a replica of the objective row, so the entering-column scan needs no exchange
and all ranks agree on it bit for bit.
"""
import numpy as np

from . import _ffi


def partition(m, world, rank):
    """Rows [begin, end) of rank `rank` -- pure arithmetic, mirrors b200lp_partition."""
    per = -(-m // max(world, 1))
    b = min(m, per * rank)
    return b, min(m, b + per)


def local_block(tab, basis, world, rank):
    """The (rows_local + 1) x C block a rank uploads: its constraint rows, then the objective row."""
    m = tab.shape[0] - 1
    b, e = partition(m, world, rank)
    blk = np.ascontiguousarray(np.vstack([tab[b:e], tab[m:m + 1]]))
    return blk, np.ascontiguousarray(basis[b:e]), (b, e)


def assemble_solution(pieces, m):
    """pieces[rank] = (rhs_local[rows+1], obj_row[C], basis_local[rows]) -> (rhs[m+1], obj_row, basis).

    The objective row and the objective value are replicas; rank 0's copy is returned."""
    rhs = np.empty(m + 1)
    basis = np.empty(m, np.int32)
    off = 0
    for rhs_l, _, basis_l in pieces:
        k = len(basis_l)
        rhs[off:off + k] = rhs_l[:k]
        basis[off:off + k] = basis_l
        off += k
    assert off == m, "shards do not cover the constraint rows"
    rhs[m] = pieces[0][0][-1]
    return rhs, np.array(pieces[0][1]), basis


def _broadcast_unique_id(group, rank, make_id):
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def solve_sharded(tab, basis, is_max=True, opts=None, group=None, device_factory=None,
                  unique_id_factory=None):
    """Solve one LP across all ranks of `group`; every rank passes the same full host tableau.

    Returns (status, iterations, rhs[m+1], obj_row[C], basis[m]) on every rank; `tab`'s RHS
    column, objective row and `basis` are updated in place like b200lp_solve does.
    device_factory / unique_id_factory exist so the host logic can be exercised on CPU (gloo)
    with a stand-in device; by default they are the CUDA library and fail without a GPU."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device_factory = device_factory or _ffi.DeviceTableau
    unique_id_factory = unique_id_factory or _ffi.comm_unique_id
    R, C = tab.shape
    m = R - 1
    uid = _broadcast_unique_id(group, rank, unique_id_factory)
    blk, blk_basis, _ = local_block(tab, basis, world, rank)
    dev = device_factory(R, C, is_max, opts, shard=(rank, world, uid))
    try:
        dev.upload(blk, blk_basis)
        status, res, _ = dev.iterate(0)
        piece = dev.download_solution()
    finally:
        dev.close()
    pieces = [None] * world
    dist.all_gather_object(pieces, tuple(np.asarray(x) for x in piece), group=group)
    rhs, obj_row, full_basis = assemble_solution(pieces, m)
    tab[:, C - 1] = rhs
    tab[m, :] = obj_row
    basis[:] = full_basis
    return status, int(res.iterations), rhs, obj_row, full_basis
