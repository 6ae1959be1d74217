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
def host_cores():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm asks for all of them itself)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def workload_string(workload):
    nx, ny, nz, k, P, radius = WORKLOADS[workload]
    return (f"{workload}: LETKF {nx}x{ny}x{nz}, {k} members, {P} obs, radius {radius}, canonical (Gaspari-Cohn R-localisation, "
            "symmetric square-root transform, X'W), inflation 1.0")


def cpu_sample(workload, seconds_hint=8.0, nthreads=0):
    """Times the CPU arm (oracle/cpu_baseline.c: the same canonical LETKF with a host cell index, tridiagonal-QL
    eigensolver, OpenMP over columns, -O3 -march=native, built on this machine) on a bounded tile of the workload: same
    k, ALL levels, same observation density, radius and localisation.  Returns columns/s and a description."""
    from metada_b200 import synthetic as syn
    from oracle import orc
    nx, ny, nz, k, P, radius = WORKLOADS[workload]
    threads = nthreads if nthreads > 0 else host_cores()
    per_col = 2.8e-9 * k ** 3 + 6e-9 * nz * k * k          # ~1.5 ms per k = 80, 60-level column per core
    ncol_target = max(256, int(seconds_hint * threads / per_col))
    t = int(min(min(nx, ny), max(16, round(ncol_target ** 0.5))))
    while t > 16 and 8.0 * t * t * nz * k > 3.0e9:           # keep the tile below 3 GB of host memory
        t -= 8
    dens = P / float(nx * ny)
    Pt = max(1, int(round(dens * t * t)))
    # same statistics as the device generator (smooth O(1) mean + 0.5 N(0,1)), drawn with NumPy's generator: the
    # bit-identical hash generator costs ~10x the timed region at this size and the CPU arm needs no bit parity
    rng = np.random.default_rng(1000)
    lev, gj, gi = np.meshgrid(np.arange(nz), np.arange(t), np.arange(t), indexing="ij")
    mean = syn.truth(gi, gj, lev, t, t)
    X = np.empty((k, nz, t, t))
    for m in range(k):
        X[m] = mean + 0.5 * rng.standard_normal(mean.shape)
    o = syn.observations(Pt, t, t, nz, seed=42, sigma=SIGMA)
    t0 = time.perf_counter()
    _, tot = orc.cpu_baseline_letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=radius,
                                    inflation=INFLATION, nthreads=threads, inplace=True)
    dt = time.perf_counter() - t0
    cols = t * t
    pbar = tot / cols
    gf_core = flops_per_column(k, pbar, nz) * cols / dt / threads / 1e9
    return {"value": cols / dt, "unit": "columns/s", "cores": threads, "kind": "port",
            "gflops_per_core": gf_core,
            "sample": f"{t}x{t}-column tile x all {nz} levels of {workload} (k={k}, {Pt} obs at the same density, radius {radius}, "
                      f"canonical / Gaspari-Cohn), oracle/cpu_baseline.c (cell index, tridiagonal QL, OpenMP, -O3 -march=native) on "
                      f"{threads} threads, {dt:.2f} s, mean p_loc {pbar:.1f} (tile edges see fewer observations than the full "
                      f"grid: favours the CPU), {gf_core:.2f} credited GFLOP/s per core",
            "seconds": dt}


def cpu_ref_as_written_c1():
    """BASELINE.md section 2 `ref-as-written` at C1 (100x100, 20 members, 1e3 obs, radius 10): LETKF.hpp:197-206 calls
    obs_op.apply(member, obs) -- H over ALL P observations -- once per (local observation, member) at every grid point.
    The count of those calls is exact (sum of p_loc x k); one call is timed here with the oracle's H (no allocations;
    the reference's own apply measured 175 ns per observation in the survey, ~9x this port's); the per-point algebra
    is timed by the oracle's as-written emulator on a 20x20 sub-grid.  Estimate = calls x time per call + algebra."""
    from metada_b200 import synthetic as syn
    from oracle import orc
    nx = ny = 100
    k, P, radius = 20, 1000, 10.0
    X = syn.ensemble(k, nx, ny, 1, seed=1000)
    o = syn.observations(P, nx, ny, 1, seed=42, sigma=SIGMA)
    counts = orc.select_counts(nx, ny, o["x"], o["y"], radius)
    calls = int(counts.sum()) * k
    t0 = time.perf_counter()
    reps = 2000
    for _ in range(reps):
        orc.hx_idw4(X[0], o["x"], o["y"], o["z"], o["valid"])
    t_call = (time.perf_counter() - t0) / reps
    t, Pt = 20, 40
    Xs = syn.ensemble(k, t, t, 1, seed=1000)
    os_ = syn.observations(Pt, t, t, 1, seed=42, sigma=SIGMA)
    t0 = time.perf_counter()
    orc.letkf(Xs, os_["x"], os_["y"], os_["z"], os_["value"], os_["err"], os_["valid"], radius=radius, mode=orc.MODE_REF_COMPAT,
              loc=orc.LOC_CUTOFF, semantics=orc.SEM_AS_WRITTEN, nthreads=1)
    t_alg = (time.perf_counter() - t0) * (nx * ny) / (t * t)
    est = calls * t_call + t_alg
    return {"columns_per_s_estimated": nx * ny / est, "seconds_estimated": est, "cores": 1, "apply_calls": calls,
            "seconds_per_apply_call_port": t_call, "algebra_seconds_scaled_from_20x20": t_alg,
            "seconds_estimated_with_reference_apply_cost": calls * P * 175e-9 + t_alg,
            "note": "exact call count x measured cost of one H pass over the 1e3 observations; the last figure uses the "
                    "reference's own IdentityObsOperator::apply cost measured by the survey (175 ns per observation)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(3.0, min(12.0, 60.0 / max(1, args.steps + args.warmup)))
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
            "config": {"workload": workload_string(args.workload)},
            "cpu_baseline": {"value": v, "unit": "columns/s", "cores": last["cores"], "kind": "port", "sample": last["sample"],
                             "gflops_per_core": last["gflops_per_core"]},
            "e2e": {"value": v, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's own LETKF.hpp needs Eigen (absent here), is single-threaded and re-applies H inside the "
                    "grid loop; this arm times a performance-minded port of the same analysis (snapshot semantics, H hoisted, "
                    "cell index) on all host cores -- each step is a bounded tile of the workload"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ other BASELINE configs
def measure_configs(ctx, mb, capi, syn, fp64_peak):
    """BASELINE.json configs[0..3] (and a conditioning-cliff variant of C5) on this GPU, device-timed through the C ABI:
    one warm-up + best of two analyses each, the background refilled on the device before every analysis."""
    out = {}

    def letkf_case(name, nx, ny, nz, k, P, radius, radius_v=0.0, sigma=SIGMA, mode=None, note=None):
        mode = mb.MODE_CANONICAL if mode is None else mode
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        o = syn.observations(P, nx, ny, nz, seed=42, sigma=sigma)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        prm = capi.make_params(radius, INFLATION, mode, mb.LOC_GASPARI_COHN if mode == mb.MODE_CANONICAL else mb.LOC_CUTOFF,
                               radius_v=radius_v)
        best = None
        for _ in range(3):
            ens.fill_synthetic(1000)
            st = capi.letkf_analyse(ens, obs, prm)
            if _ > 0 and (best is None or st["ms_total"] < best["ms_total"]):
                best = st
        G = nx * ny
        ntr = nz if radius_v > 0 else 1
        pbar = best["sum_local_obs"] / best["columns"]
        r = {"workload": f"LETKF {nx}x{ny}x{nz}, {k} members, {P} obs, radius {radius}" + (f", vertical radius {radius_v} levels "
                         f"({nz} transforms per column)" if radius_v > 0 else "") + (f", obs error {sigma}" if sigma != SIGMA else ""),
             "ms": best["ms_total"], "columns_per_s": G / (best["ms_total"] * 1e-3), "mean_local_obs": pbar,
             "redo_transforms": best["redo_transforms"], "small_transforms": best["small_transforms"],
             "numeric_failures": best["numeric_failures"], "state_gb": G * nz * k * 8 / 1e9}
        if mode == mb.MODE_CANONICAL:
            if ntr == 1:
                F = flops_per_column(k, pbar, nz)
                r["flops_per_column_credited"] = F
                r["roofline_frac_fp64"] = F * r["columns_per_s"] / 1e12 / fp64_peak
            else:
                # one transform per level: SURVEY 8d's k-space count (11 k^3 per transform) does not describe what runs --
                # with ~10 local observations per transform almost all are solved in observation space (p x p Jacobi,
                # letkf_smallp.cuh) -- so no FP64 fraction is claimed; transforms/s is the honest rate
                r["transforms_per_s"] = ntr * r["columns_per_s"]
                r["roofline_frac_fp64"] = None
        if note:
            r["note"] = note
        out[name] = r
        ens.close(); obs.close()

    letkf_case("C1", 100, 100, 1, 20, 1000, 10.0, note="k < 24: Jacobi column kernel")
    letkf_case("C1_ref_compat", 100, 100, 1, 20, 1000, 10.0, mode=mb.MODE_REF_COMPAT, note="LETKF.hpp:209-238 arithmetic")
    letkf_case("C3", 400, 400, 50, 40, 100000, 7.0)
    letkf_case("C4", 1000, 1000, 60, 128, 500000, 8.0, radius_v=5.0)
    letkf_case("C4_horizontal_only", 1000, 1000, 60, 128, 500000, 8.0)
    letkf_case("C5q_sigma_0.01", 256, 256, 60, 80, 29127, 8.0, sigma=0.01,
               note="accurate observations (50x below the ensemble spread): condition bounds ~7e3, 20 products per column; "
                    "within the packed kernel's limit (1e5) since round 2 -- redo_transforms counts what is not")
    letkf_case("C5q_sigma_0.004", 256, 256, 60, 80, 29127, 8.0, sigma=0.004,
               note="condition bounds ~4e4, 23 products per column + refined mean update: still on the packed kernel")
    letkf_case("C5q_sigma_0.002", 256, 256, 60, 80, 29127, 8.0, sigma=0.002,
               note="the conditioning cliff: condition bounds ~1.7e5 are beyond the packed kernel's limit, the columns go "
                    "through its redo list (full-product Newton-Schulz kernel)")
    # C2: global stochastic EnKF, n = 1e5, 40 members, 1e4 distinct observations
    nx, ny, k, P = 400, 250, 40, 10000
    ens = mb.Ensemble(ctx, nx, ny, 1, k)
    o = syn.observations(P, nx, ny, 1, seed=42, sigma=SIGMA, distinct=True)
    Z = np.random.default_rng(7).standard_normal((P, k))
    best = None
    for _ in range(3):
        ens.fill_synthetic(1000)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        ctx.sync()
        ctx.timer_start()
        capi.enkf_analyse(ens, obs, 1.0, Z=Z, want_gain_stats=False)
        ms = ctx.timer_stop()
        obs.close()
        if _ > 0:
            best = ms if best is None else min(best, ms)
    out["C2_enkf"] = {"workload": "global stochastic EnKF, state n = 1e5 (400x250), 40 members, 1e4 distinct obs, supplied perturbations",
                      "ms": best, "note": "Woodbury form in ensemble space (O(P k^2), K never stored); includes the H2D copy of Z"}
    ens.close()
    return out


def e2e_c_runtime(mb, torch, dist, job, ctx, params, obs_all, steps, rank, world, local_rank):
    """The same metric end to end through the C++ runtime behind the C ABI (mdc_stream_analyse, csrc/mdc_runtime.cpp):
    this rank's rows of the members live in PINNED HOST memory; every step streams them through the device in row slabs
    (upload || H + halo + column analysis || download on three host threads), in place; with several ranks the step
    starts with the NCCL observation-halo exchange (mdc_comm_init).  Wall clock, max over ranks (ncclAllReduce)."""
    import psutil
    gnx, gny, nz, k = job.gnx, job.gny, job.nz, job.k
    n_loc = nz * job.ny_loc * gnx
    need = n_loc * k * 8
    avail = psutil.virtual_memory().available
    if need > 0.62 * avail / max(1, world):
        return {"value": None, "unit": "columns/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "skipped": f"pinned host buffers need {need/1e9:.1f} GB, {avail/1e9:.1f} GB available"}
    host = [torch.empty(n_loc, dtype=torch.float64, pin_memory=True) for _ in range(k)]
    ptrs = [t.data_ptr() for t in host]
    sl = mb.Stream(local_rank, gnx, gny, nz, k, params.radius, row_range=(job.y0, job.y1), slab_rows=0, slots=4)
    if world > 1:
        uid = [mb.Stream.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        sl.comm_init(uid[0], rank, world)
    times, st, tm = [], None, None
    for it in range(steps + 1):           # first pass is the warm-up
        job.ens.fill_synthetic(1000)
        job.ens.download_ptrs(0, ptrs)    # the background ensemble now lives in HOST memory
        ctx.sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        st = sl.analyse(ptrs, obs_all, params, host_row0=job.y0, host_ny=job.ny_loc)   # returns when every slab is back
        dt = sl.comm_max(time.perf_counter() - t0)
        tm = sl.timings()
        if it > 0:
            times.append(dt)
    nslab, nslots = sl.nslab, sl.nslots
    sl.close()
    del host
    tot = float(sum(times))
    own = int(np.count_nonzero((obs_all["y"] >= (job.y0 if job.y0 > 0 else -(1 << 30))) & (obs_all["y"] < (job.y1 if job.y1 < gny else (1 << 30)))))
    obs_bytes = own * (3 * 4 + 8 + 8 + 8 + 1 + 8)
    extra_rows = nslab + (2 * (job.reach + 1) if world > 1 else 0)      # slab halo rows (+ edge strips) are uploaded twice
    return {"value": gnx * gny * len(times) / tot, "unit": "columns/s",
            "h2d_bytes_per_step": int(need + obs_bytes + need / max(1, job.ny_loc) * extra_rows), "d2h_bytes_per_step": int(need),
            "ms_per_step": 1e3 * tot / len(times), "steps": len(times), "slabs_per_rank": nslab, "slots": nslots,
            "columns_checked_rank0": st["columns"], "phases_last_step_rank0": tm,
            "note": "mdc_stream_analyse (C++ runtime behind the C ABI, no Python in the loop): pinned host members -> row slabs "
                    "through upload || H + obs-halo between slabs + column analysis || download, in place; "
                    + ("NCCL observation-halo exchange between ranks first (grouped ncclSend/ncclRecv from C); " if world > 1 else "")
                    + "host wall clock, max over ranks; byte counts are this rank's"}


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
    configs, configs_clocks, fp64_peak0 = None, None, None
    if world == 1 and not args.no_configs:
        # the other BASELINE configurations first (they need the device memory the C5 state takes afterwards)
        fp64_peak0 = ctx.bench_fp64_fma()
        cs = ClockSampler(local_rank)
        cs.start()
        configs = measure_configs(ctx, mb, capi, syn, fp64_peak0)
        configs_clocks = cs.stop()
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
    e2e = e2e_c_runtime(mb, torch, dist, job, ctx, params, obs_all, max(1, min(2, args.steps)), rank, world, local_rank)

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
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            if tr["kernel"].startswith("letkf_ns") and 24 <= k <= 80 and args.solver != "jacobi" and k == 80 and nz == 60:
                traffic = tr["dram_bytes_per_column"] * (G / world)
                traffic_src = tr["source"] + "; per-column figure scaled to this launch's columns"
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "LETKF analysed grid-columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.workload),
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
        if configs is not None:
            line["configs"] = configs
            line["configs_clocks"] = configs_clocks
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {kk: v for kk, v in cpu_sample(args.workload, seconds_hint=8.0).items() if kk != "seconds"}
            try:
                line["cpu_ref_as_written_c1"] = cpu_ref_as_written_c1()
            except Exception as e:  # noqa: BLE001
                line["cpu_ref_as_written_c1"] = {"error": str(e)}
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
    ap.add_argument("--no-configs", action="store_true", help="skip the block of the other BASELINE configurations (N = 1)")
    ap.add_argument("--solver", default="auto", choices=["auto", "jacobi", "ns"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
