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
    "cfg5": dict(m=4096, n=4096, degenerate=True),
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
    tab, basis = synthetic.dense_tableau(m, n, degenerate=cfg["degenerate"])
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

    m, n = cfg["m"], cfg["n"]
    R, C = m + 1, n + m + 1
    # pinned host buffer: the e2e call copies from it; every rank builds the same LP (seeded)
    host = torch.empty((R, C), dtype=torch.float64, pin_memory=True)
    tab = host.numpy()
    _, basis = synthetic.dense_tableau(m, n, degenerate=cfg["degenerate"], out=tab)

    shard = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_ffi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        shard = (rank, world, bytes(uid.cpu().numpy().tobytes()))
    # `--gpus N` without torchrun: one process drives N GPUs (the library shards in-process)
    inproc = args.gpus if (world == 1 and args.gpus > 1) else 1
    if inproc > torch.cuda.device_count():
        raise SystemExit(f"bench.py: --gpus {inproc} but {torch.cuda.device_count()} visible")
    devices = list(range(inproc)) if inproc > 1 else [local_rank]
    opts = _ffi.make_opts(devices=devices, time_kernels=True, pivot_variant=args.variant,
                          poll_interval=args.poll)
    dev = _ffi.DeviceTableau(R, C, True, opts, shard=shard)
    if shard is None:
        blk, blk_basis = tab, basis
    else:
        b, e = dev.row_begin, dev.row_end
        blk = np.ascontiguousarray(np.vstack([tab[b:e], tab[m:m + 1]]))
        blk_basis = np.ascontiguousarray(basis[b:e])

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
    value = steps_done / (ms_total / 1e3)
    ms_look = max_over_ranks(res.ms_look_kernel / max(res.look_kernel_launches, 1))
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

    # ---- e2e: through the reference-facing call, host buffers in, solution out -----------
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        if shard is None:
            eb = basis.copy()
            st2, r2, _ = _ffi.solve(tab, eb, True, _ffi.make_opts(devices=devices,
                                                                   pivot_variant=args.variant,
                                                                   poll_interval=args.poll,
                                                                   max_iters=args.e2e_max_iters))
            iters = int(r2.iterations)
            h2d, d2h = int(r2.h2d_bytes), int(r2.d2h_bytes)
            breakdown = {"ms_h2d": r2.ms_h2d, "ms_solve": r2.ms_solve, "ms_d2h": r2.ms_d2h,
                         "ms_total_in_call": r2.ms_total}
        else:
            breakdown = None
            dev.upload(blk, blk_basis)
            st2, r2, _ = dev.iterate(args.e2e_max_iters)
            dev.download_solution()
            iters = int(r2.iterations)
            h2d = blk.nbytes + blk_basis.nbytes
            d2h = 8 * (blk.shape[0] + C) + 4 * blk_basis.size
        torch.cuda.synchronize()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": iters / wall, "unit": UNIT,
               "h2d_bytes_per_step": h2d / max(iters, 1), "d2h_bytes_per_step": d2h / max(iters, 1),
               "pivots": iters, "wall_s": wall, "status": int(st2), "breakdown": breakdown,
               "what": "one solve call: pinned host tableau H2D + all pivots + solution D2H"}
    dev.close()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores --------------
    cpu = None
    if rank == 0 and world == 1 and inproc == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
        tab2, basis2 = synthetic.dense_tableau(m, n, degenerate=cfg["degenerate"], out=tab)
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world * inproc,
            "steps": steps_done,
            "warmup": args.warmup, "ms_per_step": ms_total / max(steps_done, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "m": m, "n": n, "R": R, "C": C,
                       "tableau_bytes": 8 * R * C, "sharding": f"row-block x{world * inproc}" + (" (one process)" if inproc > 1 else ""),
                       "exchange": {0: "none (one shard)", 1: "NCCL all-gather between kernels",
                                    2: "peer-mapped buffers, inside the iteration kernel"}[
                                        int(res.exchange_mode)],
                       "l2": "tableau >> 126 MB L2, no flush needed" if 8 * R * C // (world * inproc) > 3e8
                             else "tableau per GPU may be L2 resident",
                       "status_after_steps": int(st)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": load_traffic(workload) if world * inproc == 1 else None,
                         "kernel": "k_iter (rank-1 update tiles + lookahead CTAs)" if int(res.exchange_mode) != 1 else "k_update", "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": ms_pivot,
                         "ms_per_launch_isolated": ms_pivot_isolated,
                         "how": "achieved = bytes_per_launch / (CUDA-event time of the K back-to-back "
                                "launches / K); isolated = events around each launch in a second pass",
                         "peak_source": peak_src,
                         "frac_of_nominal_8TBps": achieved / 8000.0},
            "overlapped": {"what": "lookahead CTAs (entering column, ratio test, pivot-row scaling and, "
                                   "sharded, the candidate exchange) run inside the same launch, "
                                   "concurrently with the update tiles",
                           "ms_look": ms_look, "ms_exchange_nccl_fallback": ms_exch},
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
