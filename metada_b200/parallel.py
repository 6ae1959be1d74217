"""Column sharding of the LETKF across the GPUs of one node (host-side plumbing only).

LETKF columns are independent once Y' = H(X) - mean is known (snapshot semantics), so the column
grid is cut into row slabs, one per rank / GPU.  Each rank
  1. holds its slab of the ensemble (+ one read-only halo row above it: the 4-point IDW stencil of
     IdentityObsOperator.hpp:594-638 reaches one row up),
  2. applies H to the observations lying in its slab (mdc_hx_idw4),
  3. exchanges the Y' rows of observations within `radius` of a slab edge with its neighbours
     (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests) -- the only collective,
  4. analyses its own columns (mdc_letkf_analyse) with no further communication.
Observation rows carry their GLOBAL id, and the bucket index orders candidates by (cell, global id)
on a global cell grid, so every column sees the same rows in the same order whatever the rank count.

The reference has no parallelism of any kind; this module is new plumbing
around the C ABI, not a port.
"""
from __future__ import annotations

import math
import time

import numpy as np


def slab_bounds(gny: int, rank: int, world: int):
    return (gny * rank) // world, (gny * (rank + 1)) // world


def owner_of_row(y, gny: int, world: int):
    """Rank whose slab contains (clamped) row y -- inverse of slab_bounds."""
    y = np.clip(np.asarray(y, dtype=np.int64), 0, gny - 1)
    r = (y * world) // gny
    # integer slab edges are floor(gny*r/world): fix the rare off-by-one
    lo = (gny * r) // world
    hi = (gny * (r + 1)) // world
    r = np.where(y < lo, r - 1, np.where(y >= hi, r + 1, r))
    return r


def halo_plan(gny: int, world: int, reach: int):
    """For every ordered pair (src, dst), the row interval [lo, hi) of src's OWN rows whose
    observations dst needs: rows within `reach` of dst's slab.  Returns {(src, dst): (lo, hi)}."""
    plan = {}
    for dst in range(world):
        d0, d1 = slab_bounds(gny, dst, world)
        need_lo, need_hi = d0 - reach, d1 + reach   # obs rows y in [need_lo, need_hi)
        for src in range(world):
            if src == dst:
                continue
            s0, s1 = slab_bounds(gny, src, world)
            # src also owns out-of-grid obs clamped into its edge slab
            lo = max(s0 if src > 0 else -(1 << 30), need_lo)
            hi = min(s1 if src < world - 1 else (1 << 30), need_hi)
            if lo < hi:
                plan[(src, dst)] = (int(lo), int(hi))
    return plan


def select_own(obs: dict, gny: int, rank: int, world: int):
    """Indices (= global ids) of the observations owned by `rank`."""
    return np.nonzero(owner_of_row(obs["y"], gny, world) == rank)[0]


def exchange_rows(dist, rank, world, send: dict, row_doubles: int, device, as_dict=False):
    """send: {dst: tensor [n, row_doubles]} -> returns list of received tensors (by src order), or
    {src: tensor} with as_dict.  Works with any torch.distributed backend (gloo tensors on cpu, nccl
    tensors on cuda)."""
    import torch
    counts_out = torch.zeros(world, dtype=torch.int64, device=device)
    for dst, t in send.items():
        counts_out[dst] = t.shape[0]
    counts_in = torch.zeros(world, dtype=torch.int64, device=device)
    # all_to_all_single is not available on gloo; all_gather a [world] vector per rank instead
    gathered = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(gathered, counts_out)
    for src in range(world):
        counts_in[src] = gathered[src][rank]
    ops, recv = [], {}
    for src in range(world):
        n = int(counts_in[src])
        if src != rank and n > 0:
            recv[src] = torch.empty((n, row_doubles), dtype=torch.float64, device=device)
            ops.append(dist.P2POp(dist.irecv, recv[src], src))
    for dst, t in send.items():
        if t.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, t, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if as_dict:
        return recv
    return [recv[s] for s in sorted(recv)]


class SlabLetkf:
    """One rank's share of a column-sharded LETKF (world == 1: the whole domain)."""

    def __init__(self, ctx, gnx, gny, nz, k, rank=0, world=1, radius=8.0):
        import metada_b200 as mb
        self.mb, self.ctx = mb, ctx
        self.gnx, self.gny, self.nz, self.k = gnx, gny, nz, k
        self.rank, self.world = rank, world
        self.reach = int(math.floor(radius))
        self.y0, self.y1 = slab_bounds(gny, rank, world)
        self.halo_hi = 1 if self.y1 < gny else 0
        self.ny_loc = (self.y1 - self.y0) + self.halo_hi
        self.ens = mb.Ensemble(ctx, gnx, self.ny_loc, nz, k)
        self.ens.set_domain(0, self.y0, gnx, gny, gnx, self.y1 - self.y0)
        self.obs = None
        self.plan = halo_plan(gny, world, self.reach)
        self._own_y = None
        self.halo_rows_last = 0

    # ---- observations
    def set_observations(self, obs_all: dict):
        if self.obs is not None:
            self.obs.close()
        own = select_own(obs_all, self.gny, self.rank, self.world) if self.world > 1 else np.arange(len(obs_all["x"]))
        self._own_y = obs_all["y"][own]
        self.obs = self.mb.Observations(self.ctx, obs_all["x"][own], obs_all["y"][own], obs_all["z"][own],
                                        obs_all["value"][own], obs_all["err"][own], obs_all["valid"][own],
                                        gid=own.astype(np.int64))
        self.h2d_obs_bytes = int(len(own) * (3 * 4 + 8 + 8 + 8 + 1))

    # ---- analysis
    def analyse(self, params, dist=None):
        from . import capi
        if self.world > 1:
            import torch
            import torch.distributed as tdist
            dist = dist or tdist
            self.obs.hx(self.ens)
            rd = self.obs.row_doubles()
            dev = torch.device("cuda", torch.cuda.current_device())
            send = {}
            for (src, dst), (lo, hi) in self.plan.items():
                if src != self.rank:
                    continue
                n = int(np.count_nonzero((self._own_y >= lo) & (self._own_y < hi)))
                buf = torch.empty((max(n, 1), rd), dtype=torch.float64, device=dev)
                got = self.obs.pack_rows(lo, hi, buf.data_ptr(), n) if n > 0 else 0
                assert got == n, (got, n)
                send[dst] = buf[:n]
            torch.cuda.synchronize()
            recv = exchange_rows(dist, self.rank, self.world, send, rd, dev)
            torch.cuda.synchronize()
            self.halo_rows_last = 0
            for t in recv:
                self.obs.append_rows(t.data_ptr(), t.shape[0])
                self.halo_rows_last += t.shape[0]
            self.ctx.sync()
        # observations change every assimilation cycle, so the bucket index belongs to the step:
        # rebuild it even when this object is reused with the same observations (bench.py)
        if int(math.floor(params.radius)) > self.reach:
            raise ValueError(f"SlabLetkf was planned for radius < {self.reach + 1}; analyse() got {params.radius}")
        self.obs.index_build(max(1, int(math.ceil(params.radius))))
        return capi.letkf_analyse(self.ens, self.obs, params)

    # ---- end to end with host buffers
    def e2e_measure(self, params, obs_all, steps=1, dist=None):
        """Same metric through the C ABI with HOST (pinned) member buffers: every step uploads the
        rank's slab of all k members, analyses, and downloads the analysed members."""
        import psutil
        import torch
        n_loc = self.nz * self.ny_loc * self.gnx
        need = n_loc * self.k * 8
        avail = psutil.virtual_memory().available
        G = self.gnx * self.gny
        if need > 0.62 * avail / max(1, self.world if dist is not None else 1):
            return {"value": None, "unit": "columns/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "skipped": f"pinned host buffers need {need/1e9:.1f} GB, {avail/1e9:.1f} GB available"}
        host = [torch.empty(n_loc, dtype=torch.float64, pin_memory=True) for _ in range(self.k)]
        ptrs = [t.data_ptr() for t in host]
        if self.world == 1:
            return self._e2e_streamed(params, obs_all, steps, ptrs, need, host)
        if not getattr(self, "e2e_no_stream", False):
            return self._e2e_streamed_sharded(params, obs_all, steps, ptrs, need, host, dist)
        times = []
        for it in range(steps + 1):           # first pass is the warm-up
            self.ens.fill_synthetic(1000)
            self.ens.download_ptrs(0, ptrs)   # background ensemble now lives in HOST memory
            self.ctx.sync()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            self.ens.upload_ptrs(0, ptrs)
            self.set_observations(obs_all)
            self.ctx.sync()
            t1 = time.perf_counter()
            self.analyse(params, dist)
            t2 = time.perf_counter()
            self.ens.download_ptrs(0, ptrs)   # synchronises
            mean = self.ens.mean()            # analysis mean read back (LETKF.hpp:116)
            dt = time.perf_counter() - t0
            phases = {"upload_ms": 1e3 * (t1 - t0), "analyse_ms": 1e3 * (t2 - t1),
                      "download_ms": 1e3 * (t0 + dt - t2)}
            if dist is not None:
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt[0])
            if it > 0:
                times.append(dt)
            del mean
        del host
        tot = float(sum(times))
        return {"value": G * len(times) / tot, "unit": "columns/s",
                "h2d_bytes_per_step": int(need + self.h2d_obs_bytes), "d2h_bytes_per_step": int(need + n_loc * 8),
                "ms_per_step": 1e3 * tot / len(times), "steps": len(times), "phases_last_step": phases,
                "note": "pinned host members -> mdc_ens_upload_members -> mdc_letkf_analyse -> "
                        "mdc_ens_download_members + mdc_ens_mean, host wall clock, max over ranks"}

    def _e2e_streamed(self, params, obs_all, steps, ptrs, need, host):
        """Single GPU: the host members are streamed through the device in row slabs
        (metada_b200/pipeline.py): upload || analyse || download overlap on separate streams."""
        from .pipeline import StreamedLetkf
        G = self.gnx * self.gny
        sl = StreamedLetkf(self.ctx.device, self.gnx, self.gny, self.nz, self.k, params.radius,
                           slab_rows=32, slots=4)
        times = []
        for it in range(steps + 1):           # first pass is the warm-up
            self.ens.fill_synthetic(1000)
            self.ens.download_ptrs(0, ptrs)   # background ensemble now lives in HOST memory
            self.ctx.sync()
            t0 = time.perf_counter()
            st = sl.analyse(ptrs, obs_all, params)   # in place in the host members; returns when all slabs are back
            dt = time.perf_counter() - t0
            if it > 0:
                times.append(dt)
        sl.close()
        del host
        tot = float(sum(times))
        obs_bytes = int(len(obs_all["x"]) * (3 * 4 + 8 + 8 + 8 + 1))
        halo_rows_bytes = int(need / max(1, self.gny) * sl.nslab)      # one halo row per slab is uploaded twice
        return {"value": G * len(times) / tot, "unit": "columns/s",
                "h2d_bytes_per_step": int(need + obs_bytes + halo_rows_bytes), "d2h_bytes_per_step": int(need),
                "ms_per_step": 1e3 * tot / len(times), "steps": len(times), "slabs": sl.nslab, "columns_checked": st["columns"],
                "note": "pinned host members streamed in %d row slabs through a 3-stage pipeline (4 slots/streams): mdc_ens_upload_members_rows -> "
                        "mdc_hx_idw4 -> obs-halo pack/append between slabs -> mdc_letkf_analyse -> "
                        "mdc_ens_download_members_rows (in place); host wall clock" % sl.nslab}


    # ---- streamed, column-sharded: every rank streams its own row range through its GPU
    def edge_pack(self, ptrs, obs_all):
        """H(x) from the HOST members on the observations of this rank's edge strips -- the rows that
        other ranks' columns can reach (halo_plan) -- packed per destination rank.  Only the strips
        (<= reach + 1 rows each) are uploaded.  Returns {dst: tensor [n, k + 8] on the device}."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        rd = self.k + 8
        own = select_own(obs_all, self.gny, self.rank, self.world)
        ys = obs_all["y"][own]
        send = {}
        if not hasattr(self, "_strip"):
            self._strip = {}
        for side in ("top", "bottom"):
            dsts = [d for (s_, d) in self.plan if s_ == self.rank and (d < self.rank if side == "top" else d > self.rank)]
            if not dsts:
                continue
            lo = min(self.plan[(self.rank, d)][0] for d in dsts)
            hi = max(self.plan[(self.rank, d)][1] for d in dsts)
            r0, r1 = max(lo, self.y0), min(hi, self.y1)
            if r0 >= r1:
                continue
            halo = 1 if r1 < self.gny else 0
            key = (side, r0, r1)
            if key not in self._strip:
                ens = self.mb.Ensemble(self.ctx, self.gnx, (r1 - r0) + halo, self.nz, self.k)
                ens.set_domain(0, r0, self.gnx, self.gny, self.gnx, r1 - r0)
                self._strip[key] = [ens, None]
            ens, ob = self._strip[key]
            ens.upload_rows(ptrs, self.ny_loc, r0 - self.y0)
            sel = (ys >= lo) & (ys < hi)
            idx = own[sel]
            args = (obs_all["x"][idx], obs_all["y"][idx], obs_all["z"][idx], obs_all["value"][idx],
                    obs_all["err"][idx], obs_all["valid"][idx])
            if ob is None:
                ob = self._strip[key][1] = self.mb.Observations(self.ctx, *args, gid=idx.astype(np.int64))
            else:
                ob.assign(*args, gid=idx.astype(np.int64))
            ob.hx(ens)
            for d in dsts:
                l, h = self.plan[(self.rank, d)]
                n = int(np.count_nonzero((ys[sel] >= l) & (ys[sel] < h)))
                buf = torch.empty((max(n, 1), rd), dtype=torch.float64, device=dev)
                got = ob.pack_rows(l, h, buf.data_ptr(), n) if n > 0 else 0
                assert got == n, (got, n)
                send[d] = buf[:n]
        self.ctx.sync()
        return send

    def streamed_analyse(self, sl, ptrs, obs_all, params, recv: dict):
        """recv: {src rank: tensor of packed rows}.  Streams this rank's rows [y0, y1) through `sl`
        (a pipeline.StreamedLetkf built with row_range=(y0, y1)), in place in the host members."""
        top = [(t.data_ptr(), int(t.shape[0])) for src, t in sorted(recv.items()) if src < self.rank and t.shape[0] > 0]
        bot = [(t.data_ptr(), int(t.shape[0])) for src, t in sorted(recv.items()) if src > self.rank and t.shape[0] > 0]
        return sl.analyse(ptrs, obs_all, params, host_row0=self.y0, host_ny=self.ny_loc, ext_top=top, ext_bottom=bot)

    def _e2e_streamed_sharded(self, params, obs_all, steps, ptrs, need, host, dist):
        import torch
        import torch.distributed as tdist
        from .pipeline import StreamedLetkf
        dist = dist or tdist
        G = self.gnx * self.gny
        dev = torch.device("cuda", torch.cuda.current_device())
        sl = StreamedLetkf(self.ctx.device, self.gnx, self.gny, self.nz, self.k, params.radius,
                           slab_rows=32, slots=4, row_range=(self.y0, self.y1))
        times, phases = [], {}
        for it in range(steps + 1):           # first pass is the warm-up
            self.ens.fill_synthetic(1000)
            self.ens.download_ptrs(0, ptrs)   # background ensemble now lives in HOST memory
            self.ctx.sync()
            dist.barrier()
            t0 = time.perf_counter()
            send = self.edge_pack(ptrs, obs_all)
            torch.cuda.synchronize()
            recv = exchange_rows(dist, self.rank, self.world, send, self.k + 8, dev, as_dict=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            st = self.streamed_analyse(sl, ptrs, obs_all, params, recv)
            dt = time.perf_counter() - t0
            phases = {"edge_halo_ms": 1e3 * (t1 - t0), "stream_ms": 1e3 * (t0 + dt - t1)}
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if it > 0:
                times.append(float(tt[0]))
        sl.close()
        del host
        tot = float(sum(times))
        own_obs = int(np.count_nonzero(owner_of_row(obs_all["y"], self.gny, self.world) == self.rank))
        obs_bytes = own_obs * (3 * 4 + 8 + 8 + 8 + 1)
        extra_rows = sl.nslab + 2 * (self.reach + 1)        # slab halo rows + the edge strips are uploaded twice
        return {"value": G * len(times) / tot, "unit": "columns/s",
                "h2d_bytes_per_step": int(need + obs_bytes + need / max(1, self.ny_loc) * extra_rows),
                "d2h_bytes_per_step": int(need), "ms_per_step": 1e3 * tot / len(times), "steps": len(times),
                "slabs_per_rank": sl.nslab, "columns_checked_rank0": st["columns"], "phases_last_step": phases,
                "note": "per rank: edge-strip H + NCCL obs-halo exchange, then pinned host members streamed in row "
                        "slabs through a 3-stage pipeline (upload || analyse || download, 4 slots); host wall clock, "
                        "max over ranks; byte counts are this rank's"}

    def close(self):
        for ens, ob in getattr(self, "_strip", {}).values():
            ens.close()
            if ob is not None:
                ob.close()
        self._strip = {}
        if self.obs is not None:
            self.obs.close()
            self.obs = None
        self.ens.close()


# ------------------------------------------------------------------------------------------------------------------
# GEOGRAPHIC observations on a decomposed domain (the WRF-shaped case, sharded)
#
# Same row slabs.  What changes is who needs which observation: the selection is by haversine kilometres between the
# column's and the observation's latitude / longitude (Location.hpp:213-217, 349-357), which grid rows say nothing
# about on a curvilinear grid.  Every rank therefore keeps the (small) global coordinate arrays in a one-level,
# one-member store -- it serves the nearest-grid-point search of ALL observations (ownership = the slab of the located
# row, as for GRID observations) and hands each slab its window of the geography with the GLOBAL frame, so that the
# lat / lon lattice of the bucket index is the one-shot run's -- and sends rank r the own observations inside r's
# columns' bounding box widened by the radius (a cover: the haversine test of the selection decides).  Rows carry
# latitude, longitude, level and variable next to Y' (k + 12 doubles).  The result is bit-identical to the
# single-store analysis for any rank count.

def geo_reach_degrees(radius_km: float, latmin: float, latmax: float):
    """Largest latitude / longitude difference (degrees) between a column and an observation within radius_km of it
    (the bound geo_prepare_index uses, csrc/geo_api.inl)."""
    delta = max(radius_km, 0.0) / 6371.0
    phic = math.radians(max(abs(latmin), abs(latmax)))
    dlat = math.degrees(delta) * (1.0 + 1e-9) + 1e-12
    dlon = math.degrees(math.asin(min(1.0, math.sin(delta) / math.cos(phic)))) * (1.0 + 1e-9) + 1e-12
    return dlat, dlon


def geo_halo_boxes(lat, lon, frame: dict, world: int, radius_km: float):
    """Per rank: (lat_lo, lat_hi, u_lo, u_hi) = bounding box of the slab's columns in the geography's frame
    (u = longitude offset from frame['lon_c'], unwrapped), widened by the reach of the radius."""
    gny = lat.shape[0]
    dlat, dlon = geo_reach_degrees(radius_km, frame["latmin"], frame["latmax"])
    u = (lon - frame["lon_c"]) - 360.0 * np.rint((lon - frame["lon_c"]) / 360.0)
    boxes = []
    for r in range(world):
        y0, y1 = slab_bounds(gny, r, world)
        boxes.append((float(lat[y0:y1].min()) - dlat, float(lat[y0:y1].max()) + dlat,
                      float(u[y0:y1].min()) - dlon, float(u[y0:y1].max()) + dlon))
    return boxes


class GeoSlabLetkf:
    """One rank's share of a column-sharded LETKF with GEOGRAPHIC observations (all three modes)."""

    def __init__(self, ctx, lat, lon, vertical, nz, k, rank=0, world=1, radius_km=50.0, var_nlev=None):
        import metada_b200 as mb
        self.mb, self.ctx = mb, ctx
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        self.gny, self.gnx = lat.shape
        self.nz, self.k, self.rank, self.world, self.radius_km = nz, k, rank, world, float(radius_km)
        self.whole = mb.Ensemble(ctx, self.gnx, self.gny, 1, 1)      # geography only: locate + windows
        self.whole.set_geography(lat, lon, vertical)
        self.frame = self.whole.geography_frame()
        self.y0, self.y1 = slab_bounds(self.gny, rank, world)
        self.halo_hi = 1 if self.y1 < self.gny else 0
        self.ny_loc = (self.y1 - self.y0) + self.halo_hi
        self.ens = mb.Ensemble(ctx, self.gnx, self.ny_loc, nz, k)
        self.ens.set_domain(0, self.y0, self.gnx, self.gny, self.gnx, self.y1 - self.y0)
        self.ens.set_geography_from(self.whole)
        if var_nlev is not None:
            self.ens.set_variables(var_nlev)
        self.boxes = geo_halo_boxes(lat, lon, self.frame, world, self.radius_km)
        self.obs = None
        self.halo_rows_last = 0

    def set_observations(self, o: dict):
        """o: lat, lon, level, value, err, valid [, var] of ALL observations (every rank passes the same arrays)."""
        mb = self.mb
        if self.obs is not None:
            self.obs.close()
        everything = mb.Observations.geographic(self.ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
        everything.locate(self.whole)
        _, y, _ = everything.grid_coords()
        everything.close()
        own = np.nonzero(owner_of_row(y, self.gny, self.world) == self.rank)[0] if self.world > 1 else np.arange(len(y))
        self.own = own
        self.obs = mb.Observations.geographic(self.ctx, o["lat"][own], o["lon"][own], o["level"][own], o["value"][own],
                                              o["err"][own], o["valid"][own], gid=own.astype(np.int64))
        if o.get("var") is not None:
            self.obs.set_variables(np.asarray(o["var"])[own])
        self.obs.locate(self.whole)

    def pack_halo(self):
        """H on the own observations, then {dst: cuda tensor [n, k + 12]} of the rows dst's columns can reach."""
        import torch
        self.obs.hx(self.ens)
        rd = self.obs.row_doubles()
        dev = torch.device("cuda", torch.cuda.current_device())
        send = {}
        for dst in range(self.world):
            if dst == self.rank:
                continue
            box = self.boxes[dst]
            n = self.obs.pack_rows_geo(*box, self.frame["lon_c"], 0, 0)
            buf = torch.empty((max(n, 1), rd), dtype=torch.float64, device=dev)
            if n > 0:
                got = self.obs.pack_rows_geo(*box, self.frame["lon_c"], buf.data_ptr(), n)
                assert got == n, (got, n)
            send[dst] = buf[:n]
        self.ctx.sync()
        return send

    def analyse(self, params, dist=None, recv=None):
        """recv: {src: tensor} when the caller moved the rows itself (tests play the ranks in one process); otherwise
        the rows go through torch.distributed."""
        from . import capi
        if params.radius > self.radius_km:
            raise ValueError(f"GeoSlabLetkf was planned for a radius of {self.radius_km} km; analyse() got {params.radius}")
        if self.world > 1:
            import torch
            if recv is None:
                import torch.distributed as tdist
                dist = dist or tdist
                send = self.pack_halo()
                torch.cuda.synchronize()
                recv = exchange_rows(dist, self.rank, self.world, send, self.obs.row_doubles(),
                                     torch.device("cuda", torch.cuda.current_device()), as_dict=True)
                torch.cuda.synchronize()
            self.halo_rows_last = 0
            for src in sorted(recv):
                t = recv[src]
                if t.shape[0]:
                    self.obs.append_rows(t.data_ptr(), t.shape[0])
                    self.halo_rows_last += t.shape[0]
            self.ctx.sync()
        else:
            self.obs.hx(self.ens)
        return capi.letkf_analyse(self.ens, self.obs, params)

    def close(self):
        if self.obs is not None:
            self.obs.close()
        self.ens.close()
        self.whole.close()
