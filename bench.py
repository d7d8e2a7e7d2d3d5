#!/usr/bin/env python
"""bench.py -- simplex pivots/sec and pivot-update HBM GB/s vs roofline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one simplex iteration (entering-column scan + ratio test + rank-1 pivot update,
src/simplex.lisp:455-460) of the dense fp64 LP m=8192, n=16384 (BASELINE config 3; the tableau,
1.61 GB, is far larger than the 126 MB L2, so no flush is needed between steps).  The tableau is
resident in HBM when the clock starts; K steps are timed with CUDA events on the library's own
stream, max over ranks.  `e2e` is the same metric through the reference-facing call
(b200lp_solve: pinned host tableau -> H2D -> all pivots to optimality -> D2H of the solution).
For N > 1 (torchrun, one process per GPU) the same LP is row-block sharded: strong scaling.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg2": dict(m=1024, n=2048, degenerate=False),
    "cfg3": dict(m=8192, n=16384, degenerate=False),
    "cfg4": dict(m=16384, n=32768, degenerate=False),
    "cfg5": dict(m=4096, n=4096, degenerate=True, zero_frac=1 / 64),   # 64 cone rows: solves in 53 616 pivots
}
METRIC = "simplex_pivots_per_sec"
UNIT = "pivots/s"


def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload):
    """dram__bytes_read+write per k_pivot launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout_s=5.0):
        """nvidia-smi needs a moment to start; do not begin the timed region before it samples."""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.01)
        return False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, cfg, workload):
    """The reference's CPU implementation of the path.  The reference is Common Lisp and no Lisp
    exists in this image, so this times the oracle port (oracle/simplex_oracle.c, same pivot
    rule and rounding) with all the host threads OpenMP gives it."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from linear_programming_b200 import synthetic
    from oracle import oracle
    oracle.build()
    m, n = cfg["m"], cfg["n"]
    tab, basis = synthetic.dense_tableau(m, n, degenerate=cfg["degenerate"], zero_frac=cfg.get("zero_frac", 0.5))
    R, C = tab.shape
    cores = oracle.set_num_threads(len(os.sched_getaffinity(0)))
    budget_s = 150.0
    t_start = time.time()

    def pivots(k):
        done = 0
        for _ in range(k):
            j = oracle.find_entering_column(tab, True)
            if j < 0:
                break
            r = oracle.find_pivoting_row(tab, basis, j)
            if r < 0:
                break
            oracle.pivot(tab, basis, j, r, parallel=True)
            done += 1
            if time.time() - t_start > budget_s:
                break
        return done

    pivots(args.warmup)
    t0 = time.perf_counter()
    done = pivots(args.steps)
    dt = time.perf_counter() - t0
    value = done / dt if dt > 0 else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload, "m": m, "n": n, "R": R, "C": C},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{done} pivots of the full {R}x{C} tableau after "
                                   f"{args.warmup} warm-up pivots, OpenMP over rows"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Common Lisp (no Lisp in this image): oracle C port timed instead",
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------- our arm
def _sha(arr, dtype):
    return hashlib.sha256(np.ascontiguousarray(arr, dtype=dtype).tobytes()).hexdigest()


def _trace_sha(trace):
    return _sha(np.asarray(trace, dtype=np.int32).reshape(-1, 2), np.int32)


def load_final_fixture(config):
    """tests/golden/<config>_final.npz: what the ORACLE reaches at the end of the solve (generated
    by tools/make_full_goldens.py; the GPU box has neither /root/reference nor the minutes a CPU
    solve takes)."""
    path = os.path.join(ROOT, "tests", "golden", f"{config}_final.npz")
    if not os.path.exists(path):
        return None, path
    return np.load(path), path


def parity_prefix(dev, blk, blk_basis, tab, basis, k_prefix, world, all_min):
    """Checker leg, OUTSIDE every timed region: K pivots from a fresh upload on the GPU(s) against
    K pivots of the oracle on the host -- pivot trace, RHS column, objective row and basis, bit
    for bit.  Each rank checks the rows it owns plus its replica of the objective row."""
    from oracle import oracle
    oracle.build()
    oracle.set_num_threads(max(1, len(os.sched_getaffinity(0)) // max(world, 1)))
    dev.upload(blk, blk_basis)
    st, res, trace = dev.iterate(k_prefix)
    rhs, obj, g_basis = dev.download_solution()
    o_tab, o_basis = tab.copy(), basis.copy()
    ost, oit, otrace = oracle.solve(o_tab, o_basis, True, max_iters=k_prefix, parallel=True,
                                    trace_cap=k_prefix)
    b, e = dev.row_begin, dev.row_end
    m = tab.shape[0] - 1
    if dev.nranks == 1:
        b, e = 0, m
    out = {
        "prefix_pivots": int(oit),
        "prefix_trace_equal": bool(int(res.iterations) == oit and trace[:oit] == otrace and st == ost),
        "prefix_rhs_equal": bool(np.array_equal(rhs[:-1], o_tab[b:e, -1]) and rhs[-1] == o_tab[m, -1]),
        "prefix_obj_row_equal": bool(np.array_equal(obj, o_tab[m])),
        "prefix_basis_equal": bool(np.array_equal(g_basis, o_basis[b:e])),
    }
    del o_tab
    return {k: (v if k == "prefix_pivots" else all_min(v)) for k, v in out.items()}


def parity_final(fixture, status, iters, trace, rhs, obj, g_basis, b, e, all_min):
    """The e2e solve's end state against the committed oracle fixture, bit for bit."""
    m = int(fixture["m"])
    out = {
        "final_pivots_equal": bool(int(fixture["iterations"]) == iters and int(fixture["status"]) == status),
        "final_basis_equal": bool(np.array_equal(g_basis, fixture["basis"][b:e])),
        "final_rhs_equal": bool(np.array_equal(rhs[:-1], fixture["rhs"][b:e]) and rhs[-1] == fixture["rhs"][m]),
        "final_obj_row_equal": bool(np.array_equal(obj, fixture["obj_row"])),
        "final_trace_equal": bool(len(trace) == int(fixture["iterations"])
                                  and _trace_sha(trace) == str(fixture["trace_sha256"])),
    }
    return {k: all_min(v) for k, v in out.items()}


def run_b200(args, cfg, workload):
    import torch
    from linear_programming_b200 import _ffi, synthetic

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_min(flag):
        """True only if every rank says so."""
        if dist is None:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def new_uid():
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_ffi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())

    mps_info = None
    if args.mps:
        # f4: a standard LP file instead of the synthetic generator: read_mps -> build_tableau -> solve
        from linear_programming_b200 import external_formats, simplex
        t0 = time.perf_counter()
        with open(args.mps) as f:
            problem = external_formats.read_mps(f, args.mps_sense)      # an OBJSENSE record in the file wins
        t1 = time.perf_counter()
        built = simplex.build_tableau(problem, problem)
        t2 = time.perf_counter()
        if isinstance(built, list):
            raise SystemExit("bench.py --mps: this LP needs phase 1 (>= / = rows); the bench loop times "
                             "single-phase tableaus -- use linear_programming_b200.solve_problem")
        src_tab, basis = built.matrix, built.basis_columns
        is_max = built.instance_problem.type == "max"
        R, C = src_tab.shape
        m, n = R - 1, C - R
        mps_info = {"file": os.path.basename(args.mps), "read_mps_s": t1 - t0, "build_tableau_s": t2 - t1,
                    "rows": m, "columns": C - 1, "type": built.instance_problem.type}
        workload = f"MPS {os.path.basename(args.mps)} ({m} rows x {C - 1} columns)"
        host = torch.empty((R, C), dtype=torch.float64, pin_memory=True)
        tab = host.numpy()
        tab[:] = src_tab
        build_info = None
    else:
        is_max = True
        m, n = cfg["m"], cfg["n"]
        R, C = m + 1, n + m + 1
        # pinned host buffer: the e2e call copies from it; every rank builds the same LP (seeded).
        # f2: the fp64 tableau is written straight into it (no boxed intermediate, no second pass)
        host = torch.empty((R, C), dtype=torch.float64, pin_memory=True)
        tab = host.numpy()
        A, bvec, cvec = synthetic.dense_lp(m, n, degenerate=cfg["degenerate"], zero_frac=cfg.get("zero_frac", 0.5))
        t0 = time.perf_counter()
        _, basis = synthetic.tableau_from_lp(A, bvec, cvec, out=tab)
        t_build = time.perf_counter() - t0
        del A
        build_info = {"what": "build-tableau's fill (src/simplex.lisp:214-287) written as fp64 straight "
                              "into the pinned upload buffer", "seconds": t_build,
                      "bytes": 8 * R * C, "gb_per_s": 8e-9 * R * C / t_build}

    shard = (rank, world, new_uid()) if world > 1 else None
    # `--gpus N` without torchrun: one process drives N GPUs (the library shards in-process)
    inproc = args.gpus if (world == 1 and args.gpus > 1) else 1
    if inproc > torch.cuda.device_count():
        raise SystemExit(f"bench.py: --gpus {inproc} but {torch.cuda.device_count()} visible")
    devices = list(range(inproc)) if inproc > 1 else [local_rank]
    trace_cap = 1 << 15
    opts = _ffi.make_opts(devices=devices, time_kernels=True, pivot_variant=args.variant,
                          poll_interval=args.poll, trace_capacity=trace_cap)
    dev = _ffi.DeviceTableau(R, C, is_max, opts, shard=shard)
    if shard is None:
        blk, blk_basis = tab, basis
        row_b, row_e = 0, m
    else:
        row_b, row_e = dev.row_begin, dev.row_end
        blk = np.ascontiguousarray(np.vstack([tab[row_b:row_e], tab[m:m + 1]]))
        blk_basis = np.ascontiguousarray(basis[row_b:row_e])

    # ---- device-resident K steps ---------------------------------------------------------
    # Timed region: K iterations enqueued back to back (no host work, no events between launches),
    # bracketed by CUDA events on the library's stream.  A second, separate pass of K iterations
    # records events around every launch to isolate the kernel's own duration.
    dev.upload(blk, blk_basis)
    dev.set_time_kernels(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    dev.iterate(args.warmup)
    barrier()
    st, res, _ = dev.iterate(args.steps)
    barrier()
    steps_done = int(res.iterations)
    ms_total = max_over_ranks(res.ms_solve)
    launches = int(res.kernel_launches)
    bytes_per_launch = int(res.bytes_per_pivot)
    loop_mode = int(res.loop_mode)
    value = steps_done / (ms_total / 1e3)
    n_look = max(res.look_kernel_launches, 1)
    ms_look = max_over_ranks(res.ms_look_kernel / n_look)
    look_split = {k: max_over_ranks(1e3 * getattr(res, "ms_look_" + k) / n_look)
                  for k in ("wait", "ratio", "push", "peer_wait", "row")}
    ms_pivot_isolated = None
    ms_exch = 0.0
    if loop_mode == 1:
        dev.set_time_kernels(True)
        barrier()
        st_b, res_b, _ = dev.iterate(args.steps)
        ms_pivot_isolated = max_over_ranks(res_b.ms_pivot_kernel / max(res_b.pivot_kernel_launches, 1))
        ms_exch = max_over_ranks(res_b.ms_exchange / max(res_b.look_kernel_launches, 1))
        dev.set_time_kernels(False)
    clocks = sampler.stop() if rank == 0 else None   # 20 ms samples: warm-up, timed and isolated passes
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + per-launch pass (same kernel throughout)"
    # average launch duration over the timed region: K launches back to back in ms_total
    # (inter-launch gaps included, so this can only under-state the kernel)
    ms_pivot = ms_total / max(steps_done, 1)

    # ---- parity, part 1 (checker leg, untimed): a prefix against the oracle on the host --------
    parity = None
    if not args.no_parity and is_max:
        parity = parity_prefix(dev, blk, blk_basis, tab, basis, args.parity_prefix, world * inproc, all_min)

    # ---- e2e: through the reference-facing call, host buffers in, solution out -----------
    e2e = None
    fixture, fixture_path = (None, None) if args.mps else load_final_fixture(args.config)
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        if shard is None:
            eb = basis.copy()
            st2, r2, e_trace = _ffi.solve(tab, eb, is_max, _ffi.make_opts(
                devices=devices, pivot_variant=args.variant, poll_interval=args.poll,
                max_iters=args.e2e_max_iters, trace_capacity=trace_cap))
            iters = int(r2.iterations)
            h2d, d2h = int(r2.h2d_bytes), int(r2.d2h_bytes)
            breakdown = {"ms_h2d": r2.ms_h2d, "ms_solve": r2.ms_solve, "ms_d2h": r2.ms_d2h,
                         "ms_total_in_call": r2.ms_total}
            e_rhs, e_obj, e_basis = tab[:, C - 1].copy(), tab[m].copy(), eb
        else:
            breakdown = None
            dev.upload(blk, blk_basis)
            st2, r2, e_trace = dev.iterate(args.e2e_max_iters)
            e_rhs, e_obj, e_basis = dev.download_solution()
            iters = int(r2.iterations)
            h2d = blk.nbytes + blk_basis.nbytes
            d2h = 8 * (blk.shape[0] + C) + 4 * blk_basis.size + 8 * min(iters, trace_cap)
        torch.cuda.synchronize()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": iters / wall, "unit": UNIT,
               "h2d_bytes_per_step": h2d / max(iters, 1), "d2h_bytes_per_step": d2h / max(iters, 1),
               "pivots": iters, "wall_s": wall, "status": int(st2), "breakdown": breakdown,
               "objective": float(r2.objective),
               "basis_sha256_local_rows": _sha(e_basis, np.int32),
               "trace_sha256": _trace_sha(e_trace) if iters <= trace_cap else None,
               "what": "one solve call: pinned host tableau H2D + all pivots + solution D2H"}
        # ---- parity, part 2: the end state against the committed full-solve fixture ------
        if parity is not None and args.e2e_max_iters == 0:
            if fixture is not None:
                parity.update(parity_final(fixture, int(st2), iters, e_trace, e_rhs, e_obj, e_basis,
                                           row_b, row_e, all_min))
                parity["final_fixture"] = os.path.relpath(fixture_path, ROOT)
                parity["fixture_objective"] = float(fixture["objective"])
            else:
                parity["final_fixture"] = None
            parity["objective"] = float(r2.objective)
    if parity is not None:
        parity["checker"] = ("oracle/simplex_oracle.c on the host for the prefix; committed oracle "
                             "end state (tools/make_full_goldens.py) for the full solve; bit-exact compares")
    dev.close()

    # ---- BASELINE config 4 (the one the north star names for 2/4/8 GPUs), same row-block sharding
    cfg4 = None
    if args.with_cfg4 and not args.mps and args.config == "cfg3":
        m4, n4 = CONFIGS["cfg4"]["m"], CONFIGS["cfg4"]["n"]
        R4, C4 = m4 + 1, n4 + m4 + 1
        tc0 = time.perf_counter()
        if world > 1:
            b4, e4 = _ffi.partition(m4, world, rank)
            blk4, basis4 = synthetic.dense_block(m4, n4, b4, e4)
            dev4 = _ffi.DeviceTableau(R4, C4, True, _ffi.make_opts(devices=[local_rank]),
                                      shard=(rank, world, new_uid()))
        else:
            blk4, basis4 = synthetic.dense_block(m4, n4, 0, m4)
            dev4 = _ffi.DeviceTableau(R4, C4, True, _ffi.make_opts(devices=devices))
        t_gen = time.perf_counter() - tc0
        dev4.upload(blk4, basis4)
        del blk4
        dev4.iterate(args.warmup)
        barrier()
        st4, res4, _ = dev4.iterate(args.cfg4_steps)
        barrier()
        ms4 = max_over_ranks(res4.ms_solve)
        it4 = int(res4.iterations)
        obj4 = float(res4.objective)
        dev4.close()
        peak4, _ = load_peak()
        gbs4 = int(res4.bytes_per_pivot) * it4 / (ms4 * 1e-3) / 1e9
        cfg4 = {"workload": "dense fp64 LP m=16384 n=32768 (BASELINE cfg4), seed 1234",
                "value": it4 / (ms4 / 1e3), "unit": UNIT, "steps": it4, "ms_per_step": ms4 / max(it4, 1),
                "gbs_per_gpu": gbs4, "frac_of_measured_peak_per_gpu": gbs4 / peak4,
                "objective_after_steps": obj4, "setup_s": t_gen,
                "what": "same timed-region method as the headline, tableau resident in HBM"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores --------------
    cpu = None
    if rank == 0 and world == 1 and inproc == 1 and not args.no_cpu_baseline and not args.mps:
        from oracle import oracle
        oracle.build()
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
        tab2, basis2 = synthetic.dense_tableau(m, n, degenerate=cfg["degenerate"], out=tab,
                                                zero_frac=cfg.get("zero_frac", 0.5))
        n_cpu, t_budget = 0, 20.0
        for k in range(2):   # warm-up
            j = oracle.find_entering_column(tab2, True)
            oracle.pivot(tab2, basis2, j, oracle.find_pivoting_row(tab2, basis2, j), parallel=True)
        t0 = time.perf_counter()
        while n_cpu < 40 and time.perf_counter() - t0 < t_budget:
            j = oracle.find_entering_column(tab2, True)
            if j < 0:
                break
            oracle.pivot(tab2, basis2, j, oracle.find_pivoting_row(tab2, basis2, j), parallel=True)
            n_cpu += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
               "sample": f"{n_cpu} pivots of the full {R}x{C} tableau after 2 warm-up pivots "
                         f"(oracle C port, OpenMP over rows; the Lisp reference cannot run here)"}
        # the reference itself is single-threaded: the same port on one core, 3 pivots
        t0 = time.perf_counter()
        n1 = 0
        while n1 < 3:
            j = oracle.find_entering_column(tab2, True)
            if j < 0:
                break
            oracle.pivot(tab2, basis2, j, oracle.find_pivoting_row(tab2, basis2, j), parallel=False)
            n1 += 1
        cpu["single_core"] = {"value": n1 / (time.perf_counter() - t0), "unit": UNIT, "cores": 1,
                              "sample": f"{n1} pivots, same tableau, one thread"}

    if rank == 0:
        peak, peak_src = load_peak()
        achieved = bytes_per_launch / (ms_pivot * 1e-3) / 1e9
        ngpu = world * inproc
        kernel = {1: "k_iter (rank-1 update tiles + lookahead CTAs), one launch per pivot",
                  2: "k_persist (one cooperative kernel per call: look CTAs + tile CTAs)"}[loop_mode]
        if int(res.exchange_mode) == 1:
            kernel = "k_update (NCCL fallback: k_look + all-gather + k_update)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu,
            "steps": steps_done,
            "warmup": args.warmup, "ms_per_step": ms_total / max(steps_done, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if not args.mps else "file",
            "config": {"workload": workload, "m": m, "n": n, "R": R, "C": C},
            "run": {"tableau_bytes": 8 * R * C,
                    "sharding": f"row-block x{ngpu}" + (" (one process)" if inproc > 1 else ""),
                    "exchange": {0: "none (one shard)", 1: "NCCL all-gather between kernels",
                                 2: "peer-mapped buffers, inside the iteration kernel"}[
                                     int(res.exchange_mode)],
                    "loop": {1: "one k_iter launch per pivot (PDL)", 2: "persistent cooperative kernel"}[loop_mode],
                    "l2": "tableau >> 126 MB L2, no flush needed" if 8 * R * C // ngpu > 3e8
                          else "tableau per GPU may be L2 resident",
                    "status_after_steps": int(st)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": load_traffic(workload) if ngpu == 1 else None,
                         "kernel": kernel, "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": ms_pivot,
                         "ms_per_launch_isolated": ms_pivot_isolated,
                         "how": "achieved = algorithmic bytes per pivot (16*R_local*C) / (CUDA-event time of "
                                "the K back-to-back pivots / K); isolated = events around each launch in a "
                                "second pass (per-pivot loop only)",
                         "peak_source": peak_src,
                         "frac_of_nominal_8TBps": achieved / 8000.0},
            "overlapped": {"what": "lookahead CTAs (entering column, ratio test, pivot-row scaling and, "
                                   "sharded, the candidate exchange) run inside the same launch, "
                                   "concurrently with the update tiles",
                           "ms_look": ms_look, "look_ctas": int(res.look_ctas),
                           "us_look_split_max_over_ranks": look_split,
                           "ms_exchange_nccl_fallback": ms_exch},
            "parity": parity, "cfg4": cfg4, "build_tableau": build_info, "mps": mps_info,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--variant", type=int, default=0, help="pivot kernel variant (0 = default)")
    ap.add_argument("--poll", type=int, default=0)
    ap.add_argument("--e2e-max-iters", type=int, default=0, help="0 = solve to optimality")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle / fixture checker legs")
    ap.add_argument("--parity-prefix", type=int, default=10, help="pivots compared with the oracle")
    ap.add_argument("--no-cfg4", dest="with_cfg4", action="store_false",
                    help="skip the BASELINE config 4 sub-record")
    ap.add_argument("--cfg4-steps", type=int, default=200)
    ap.add_argument("--mps", default=None, help="bench an LP read from an MPS file (single-phase LPs)")
    ap.add_argument("--mps-sense", default="max", choices=["max", "min"],
                    help="problem type when the file has no OBJSENSE record (read-mps's second argument)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    workload = (f"dense fp64 LP m={cfg['m']} n={cfg['n']} (BASELINE {args.config}), "
                f"seed 1234{', degenerate' if cfg['degenerate'] else ''}")
    if args.impl == "reference":
        return run_reference(args, cfg, workload)
    return run_b200(args, cfg, workload)


if __name__ == "__main__":
    sys.exit(main())
