#!/usr/bin/env python
"""bench.py -- LETKF analysed grid-columns/s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU arm: the oracle port on host cores
    torchrun ... bench.py --gpus N ...                     # one rank per GPU, row-slab column sharding

A "step" is one complete analysis pass (H(x) -> Y', bucket index, per-column transform + update)
over the workload's synthetic background ensemble.  The background is regenerated on the device
before every step (the analysis is in place); that refill is input preparation and is outside the
per-step CUDA-event brackets.  Workload C5 = BASELINE.json configs[4] / the config the metric is
quoted on (1500x1500x60, 80 members, 1e6 obs): 86.4 GB of state, >> the 126 MB L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: nx, ny, nz, k, P, radius   
    "C1": (100, 100, 1, 20, 1000, 10.0),
    "C3": (400, 400, 50, 40, 100000, 7.0),
    "C5": (1500, 1500, 60, 80, 1000000, 8.0),
    "C5q": (256, 256, 60, 80, 29127, 8.0),     # C5 physics on a small grid (quick checks)
}
SIGMA = 0.1
INFLATION = 1.0


def flops_per_column(k, pbar, L):
    """SURVEY.md section 8d, canonical mode (only 9k^3 credited for the eigensolve)."""
    return k * (k + 1) * pbar + 2 * k * pbar + 9 * k ** 3 + 2 * k ** 3 + 2 * k ** 2 + 2 * L * k ** 2 + 2 * L * k


def bytes_per_column(k, L, P, G):
    return 2 * L * k * 8 + P * (k + 4) * 8 / G


def peaks():
    p = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p = {"hbm_gbs": float(m["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    except Exception:  # noqa: BLE001
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self._stop, self._t = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                               "-i", str(self.device)], text=True, timeout=5)
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def slab_bounds(gny, rank, world):
    """1-D row-slab decomposition of the column grid."""
    y0 = (gny * rank) // world
    y1 = (gny * (rank + 1)) // world
    return y0, y1


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(workload, seconds_hint=12.0, nthreads=0):
    """Times the oracle port (canonical LETKF, OpenMP over columns, brute-force local-obs scan as in
    LETKF.hpp:159-165) on a bounded tile of the workload with the same k / levels / obs density /
    radius.  Returns columns/s and a description of the sample."""
    from metada_b200 import synthetic as syn
    from oracle import orc
    nx, ny, nz, k, P, radius = WORKLOADS[workload]
    threads = nthreads if nthreads > 0 else orc.max_threads()
    # ~30 ms per k=80 column per core; size the tile for ~seconds_hint of work on `threads` cores
    per_col = 2.0e-8 * k ** 3
    ncol_target = max(64, int(seconds_hint * threads / per_col))
    t = int(min(min(nx, ny), max(8, round(ncol_target ** 0.5))))
    lev = min(nz, 8)            # levels only scale the (cheap) update; keep the tile small in RAM
    dens = P / float(nx * ny)
    Pt = max(1, int(round(dens * t * t)))
    X = syn.ensemble(k, t, t, lev, seed=1000)
    o = syn.observations(Pt, t, t, lev, seed=42, sigma=SIGMA)
    t0 = time.perf_counter()
    r = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=radius,
                  inflation=INFLATION, mode=orc.MODE_CANONICAL, loc=orc.LOC_GASPARI_COHN, nthreads=threads)
    dt = time.perf_counter() - t0
    cols = t * t
    return {"value": cols / dt, "unit": "columns/s", "cores": threads, "kind": "port",
            "sample": f"{t}x{t}-column tile x {lev} levels of {workload} (k={k}, {Pt} obs at the same density, "
                      f"radius {radius}, canonical/Gaspari-Cohn), oracle port with OpenMP, {dt:.2f} s; "
                      f"mean p_loc {float(r['counts'].mean()):.1f}; edge columns see fewer obs, "
                      f"brute-force scan is over {Pt} obs not {P} (both favour the CPU)",
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, nz, k, P, radius = WORKLOADS[args.workload]
    per_step = max(4.0, min(25.0, 100.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(args.workload, seconds_hint=per_step)
    tot_cols, tot_t, last = 0.0, 0.0, None
    for _ in range(args.steps):
        last = cpu_sample(args.workload, seconds_hint=per_step)
        tot_t += last["seconds"]
        tot_cols += last["value"] * last["seconds"]
    v = tot_cols / tot_t
    line = {"impl": "reference", "metric": "LETKF analysed grid-columns/sec", "value": v, "unit": "columns/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: LETKF {nx}x{ny}x{nz}, {k} members, {P} obs, radius {radius}, "
                                   "canonical (Gaspari-Cohn R-localisation, symmetric sqrt, X'W)"},
            "cpu_baseline": {"value": v, "unit": "columns/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's own LETKF.hpp needs Eigen (absent here) and is single-threaded; this arm times the "
                    "oracle port of the same path (snapshot semantics, H hoisted) on all host cores"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import metada_b200 as mb
    from metada_b200 import capi, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    nx, ny, nz, k, P, radius = WORKLOADS[args.workload]
    G = nx * ny
    ctx = mb.Context(local_rank)
    from metada_b200.parallel import SlabLetkf  # row-slab sharding + obs halo exchange
    job = SlabLetkf(ctx, nx, ny, nz, k, rank, world, radius)
    obs_all = syn.observations(P, nx, ny, nz, seed=42, sigma=SIGMA)
    solver = {"auto": mb.SOLVER_AUTO, "jacobi": mb.SOLVER_JACOBI, "ns": mb.SOLVER_NEWTON_SCHULZ}[args.solver]
    params = capi.make_params(radius, INFLATION, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, solver=solver)
    ns_name = "Newton-Schulz symmetric square root on FP64 DMMA, packed symmetric tiles (24<=k<=128; Jacobi otherwise)"
    solver_name = {"auto": ns_name if 24 <= k <= 128 else "one-sided block Jacobi eigen-decomposition",
                   "jacobi": "one-sided block Jacobi eigen-decomposition", "ns": ns_name}[args.solver]

    job.set_observations(obs_all)     # device SoA + H/Y' buffers are allocated once and reused

    def one_step(timed):
        job.ens.fill_synthetic(1000)
        ctx.sync()
        if world > 1:
            dist.barrier()
        ctx.timer_start()
        st = job.analyse(params)
        ms = ctx.timer_stop()
        return ms, st

    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    l0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    step_ms, stats = [], None
    for _ in range(args.steps):
        ms, stats = one_step(True)
        step_ms.append(ms)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    tot_ms = float(sum(step_ms))
    col_ms = float(stats["ms_columns"])
    red = torch.tensor([tot_ms, col_ms, float(stats["sum_local_obs"]), float(stats["columns"]),
                        float(stats["sum_sweeps"])], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = red.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tot_ms, col_ms = float(mx[0]), float(mx[1])
        sum_ploc, ncols, sum_sw = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        sum_ploc, ncols, sum_sw = float(red[2]), float(red[3]), float(red[4])
    value = G * args.steps / (tot_ms * 1e-3)
    pbar = sum_ploc / ncols

    # ---- end-to-end through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e = job.e2e_measure(params, obs_all, steps=max(1, min(2, args.steps)), dist=dist if world > 1 else None)

    if rank == 0:
        pk = peaks()
        fp64_peak = ctx.bench_fp64_fma()
        dmma_peak = ctx.bench_fp64_dmma()
        F = flops_per_column(k, pbar, nz)
        Bc = bytes_per_column(k, nz, P, G)
        cols_per_s_kernel = (G / world) / (col_ms * 1e-3)   # columns one GPU's kernel launch processes
        ach_tf = F * cols_per_s_kernel / 1e12
        ach_gb = Bc * cols_per_s_kernel / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per column from the committed ncu capture, scaled to this launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
            if tr["kernel"].startswith("letkf_ns") and 24 <= k <= 80 and args.solver != "jacobi" and k == 80 and nz == 60:
                traffic = tr["dram_bytes_per_column"] * (G / world)
                traffic_src = tr["source"] + "; per-column figure scaled to this launch's columns"
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "LETKF analysed grid-columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: LETKF {nx}x{ny}x{nz}, {k} members, {P} obs, radius {radius}, "
                                   "canonical (Gaspari-Cohn R-localisation, symmetric square-root transform, X'W), inflation 1.0",
                       "solver": solver_name,
                       "parallelism": f"row-slab column sharding x{world}, NCCL obs-halo exchange" if world > 1 else "single GPU",
                       "l2": "state (%.1f GB) >> 126 MB L2; background regenerated on device before every step" % (G * nz * k * 8 / 1e9),
                       "mean_local_obs": pbar, "mean_solver_iterations": sum_sw / ncols},
            "roofline": {"bound": "tensor", "bound_detail": "FP64 tensor path (DMMA mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind), "
                                                           "denominator = the higher measured FP64 FMA-pipe peak",
                         "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": Bc * (G / world), "kernel": "letkf_nsp_kernel" if (24 <= k <= 128 and args.solver != "jacobi") else "letkf_canonical_kernel",
                         "peak_source": "FP64 FMA microbenchmark run in this process (mdc_bench_fp64_fma); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "pipe": "FP64 tensor path (DMMA mma.sync.m8n8k4.f64) for the Newton-Schulz products, SYRK and update",
                         "fp64_dmma_peak": dmma_peak, "frac_of_dmma_peak": ach_tf / dmma_peak,
                         "flops_per_column": F, "bytes_per_column": Bc,
                         "hbm": {"achieved": ach_gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_gb / pk["hbm_gbs"], "peak_source": pk["source"]}},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "phases_ms_last_step": {kk: stats[kk] for kk in ("ms_hx", "ms_index", "ms_columns", "ms_total")},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {kk: v for kk, v in cpu_sample(args.workload, seconds_hint=12.0).items() if kk != "seconds"}
        print(json.dumps(line), flush=True)
    job.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MDC_BENCH_WORKLOAD", "C5"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--solver", default="auto", choices=["auto", "jacobi", "ns"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
