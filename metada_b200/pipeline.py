"""Streamed LETKF for ensembles that live in HOST memory (the reference's situation: every member is
a host array, `State::getDataPtr<double>()`, State.hpp:229-242).

The column grid is cut into row slabs that flow through the GPU as a software pipeline,
    upload slab s+1 (H2D)  ||  analyse slab s (kernels)  ||  download slab s-1 (D2H),
each slab on its own library context (= its own CUDA stream) driven by its own host thread, so the
two PCIe directions and the SMs are busy at the same time and the device only ever holds a few
slabs.  A slab needs (i) one read-only halo row above it for the 4-point IDW stencil of H, uploaded
with it, and (ii) the Y' rows of observations within `radius` rows of its edges, which its
neighbours pack right after their own H(x) (`mdc_obs_pack_rows`) and it appends
(`mdc_obs_append_rows`) -- the same halo mechanism as the multi-GPU sharding in parallel.py, here
between slabs on one device.  Results are bit-identical to the one-shot analysis (candidates are
ordered by (global cell, global observation id)).

Host-side plumbing only; every byte of arithmetic happens in the C-ABI library.
"""
from __future__ import annotations

import math
import threading

import numpy as np

from . import capi
from .parallel import halo_plan, owner_of_row, slab_bounds


class StreamedLetkf:
    def __init__(self, device, gnx, gny, nz, k, radius, slab_rows=32, workers=4, y_range=None):
        """y_range=(y0, y1): analyse only these rows (one rank's share in a multi-GPU run); the
        host arrays are always the full [lev][gny][gnx] members."""
        self.gnx, self.gny, self.nz, self.k = gnx, gny, nz, k
        self.reach = int(math.floor(radius))
        self.ya, self.yb = y_range if y_range is not None else (0, gny)
        rows = self.yb - self.ya
        self.nslab = max(1, (rows + slab_rows - 1) // slab_rows)
        self.workers = 1 if self.nslab == 1 else max(2, min(workers, self.nslab))   # >= 2: a slab waits for its upper neighbour's H
        self.ctxs = [capi.Context(device) for _ in range(self.workers)]
        self.bounds = [(self.ya + (rows * s) // self.nslab, self.ya + (rows * (s + 1)) // self.nslab)
                       for s in range(self.nslab)]
        self._ens = [None] * self.workers
        self._obs = [None] * self.workers
        self._pool = [None] * self.workers

    def close(self):
        for w in range(len(self.ctxs)):
            if self._ens[w] is not None:
                self._ens[w].close()
            if self._obs[w] is not None:
                self._obs[w].close()
            if self._pool[w]:
                self.ctxs[w].dev_free(self._pool[w][0])
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def analyse(self, member_ptrs, obs, params, extra_halo=None):
        """member_ptrs: k host pointers (pinned for full PCIe speed) to [nz][gny][gnx] float64 arrays,
        updated IN PLACE.  obs: dict of global observation arrays.  Returns summed stats.
        Device buffers (one ensemble slab, one observation store and a halo pool per worker) are
        allocated on first use and reused: cudaMalloc/cudaFree would serialise the streams."""
        import metada_b200 as mb
        nslab, R, W = self.nslab, self.reach, self.workers
        gid_all = np.arange(len(obs["y"]), dtype=np.int64)
        own_idx, halo_n = [], []
        for s, (y0, y1) in enumerate(self.bounds):
            lo = y0 if y0 > 0 else -(1 << 30)
            hi = y1 if y1 < self.gny else (1 << 30)
            idx = np.nonzero((obs["y"] >= lo) & (obs["y"] < hi))[0]
            own_idx.append(idx)
            ys = obs["y"][idx]
            cnt = {}
            for dst in (s - 1, s + 1):
                if 0 <= dst < nslab:
                    d0, d1 = self.bounds[dst]
                    cnt[dst] = (d0 - R, d1 + R, int(np.count_nonzero((ys >= d0 - R) & (ys < d1 + R))))
            halo_n.append(cnt)
        rd = self.k + 8
        max_rows = max(y1 - y0 for (y0, y1) in self.bounds) + 1
        # halo pool per worker: slots for every slab that worker will process (kept until the end)
        for w in range(W):
            need = sum(n for s in range(w, nslab, W) for (_, _, n) in halo_n[s].values()) * rd * 8
            if not self._pool[w] or self._pool[w][1] < need:
                if self._pool[w]:
                    self.ctxs[w].dev_free(self._pool[w][0])
                self._pool[w] = (self.ctxs[w].dev_malloc(max(need, 8)), need)
            if self._ens[w] is None:
                self._ens[w] = mb.Ensemble(self.ctxs[w], self.gnx, max_rows, self.nz, self.k)
        h_done = [threading.Event() for _ in range(nslab)]
        halo_buf = [dict() for _ in range(nslab)]      # slab s -> {dst: (devptr, nrows)}
        stats = [None] * nslab
        errors = []

        def run_slab(s, w, pool_off):
            ctx, ens = self.ctxs[w], self._ens[w]
            y0, y1 = self.bounds[s]
            halo_hi = 1 if y1 < self.gny else 0
            ens.set_rows((y1 - y0) + halo_hi)
            ens.set_domain(0, y0, self.gnx, self.gny, self.gnx, y1 - y0)
            ens.upload_rows(member_ptrs, self.gny, y0)
            idx = own_idx[s]
            args = (obs["x"][idx], obs["y"][idx], obs["z"][idx], obs["value"][idx], obs["err"][idx],
                    obs["valid"][idx], gid_all[idx])
            if self._obs[w] is None:
                self._obs[w] = mb.Observations(ctx, *args[:6], gid=args[6])
            else:
                self._obs[w].assign(*args[:6], gid=args[6])
            ob = self._obs[w]
            ob.hx(ens)
            for dst, (lo, hi, n) in halo_n[s].items():
                if n > 0:
                    p = self._pool[w][0] + pool_off
                    pool_off += n * rd * 8
                    got = ob.pack_rows(lo, hi, p, n)
                    assert got == n, (got, n)
                    halo_buf[s][dst] = (p, n)
            h_done[s].set()
            for src in (s - 1, s + 1):
                if 0 <= src < nslab:
                    h_done[src].wait()
                    if errors:
                        raise RuntimeError("another slab failed")
                    if s in halo_buf[src]:
                        p, n = halo_buf[src][s]
                        ob.append_rows(p, n)
            if extra_halo is not None:
                extra_halo(s, ob)
            stats[s] = capi.letkf_analyse(ens, ob, params)
            ens.download_rows(member_ptrs, self.gny, y0, y1 - y0)
            return pool_off

        def worker(w):
            try:
                off = 0
                for s in range(w, nslab, W):
                    off = run_slab(s, w, off)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                for ev in h_done:
                    ev.set()

        threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        tot = {"columns": 0, "sum_local_obs": 0, "sum_sweeps": 0, "max_local_obs": 0, "max_sweeps": 0,
               "numeric_failures": 0}
        for st in stats:
            for kk in ("columns", "sum_local_obs", "sum_sweeps", "numeric_failures"):
                tot[kk] += st[kk]
            for kk in ("max_local_obs", "max_sweeps"):
                tot[kk] = max(tot[kk], st[kk])
        return tot
