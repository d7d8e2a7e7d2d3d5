"""torchrun worker for tests/test_gpu_sharded.py: one process per GPU, NCCL, real kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from linear_programming_b200 import _ffi, sharded, synthetic
    from oracle import oracle
    m, n = int(sys.argv[1]), int(sys.argv[2])
    rule = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    degenerate = len(sys.argv) > 4 and sys.argv[4] == "degenerate"
    rank, local_rank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tab, basis = synthetic.dense_tableau(m, n, seed=77, degenerate=degenerate)
    ref_tab, ref_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(ref_tab, ref_basis, True, rule=rule, max_iters=50000,
                                    trace_cap=1 << 16, parallel=True)
    opts = _ffi.make_opts(devices=[local_rank], pivot_rule=rule, max_iters=50000,
                          trace_capacity=1 << 16)
    traces = []

    def factory(*a, **k):
        d = _ffi.DeviceTableau(*a, **k)
        orig = d.iterate

        def it(mi=0):
            out = orig(mi)
            traces.append(out[2])
            # full local block must equal the oracle's rows bit for bit
            blk, bb = d.download()
            b, e = d.row_begin, d.row_end
            assert np.array_equal(blk[:-1], ref_tab[b:e]), "constraint rows differ"
            assert np.array_equal(blk[-1], ref_tab[-1]), "objective replica differs"
            assert np.array_equal(bb, ref_basis[b:e])
            return out
        d.iterate = it
        return d

    st, it, rhs, obj, full_basis = sharded.solve_sharded(tab, basis, True, opts, device_factory=factory)
    assert st == ost, (st, ost)
    assert it == oit, (it, oit)
    assert traces[0] == otrace, "pivot sequence differs from the unsharded oracle"
    assert np.array_equal(rhs, ref_tab[:, -1]) and np.array_equal(obj, ref_tab[-1])
    assert np.array_equal(full_basis, ref_basis)
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_OK world={dist.get_world_size()} pivots={it} objective={rhs[-1]!r}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
