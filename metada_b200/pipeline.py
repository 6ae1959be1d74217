"""Streamed LETKF for ensembles that live in HOST memory (the reference's situation: every member is
a host array, `State::getDataPtr<double>()`, State.hpp:229-242).

The column grid is cut into row slabs that flow through the GPU as a three-stage software pipeline

    uploader:    H2D copy of slab s+2 + member transpose          (its slot's stream)
    compute:     H(x) of slab s+1, obs-halo append, analyse slab s (their slots' streams)
    downloader:  member transpose + D2H copy of slab s-1           (its slot's stream)

Each slab lives in one of a few reusable SLOTS (library context = CUDA stream + copy stream, ensemble
slab); three host threads hand slots to each other through queues, so both PCIe directions and the SMs
are busy at the same time and the device only ever holds a few slabs.  Every slab has its own (small)
observation store, uploaded before the member traffic starts: a small copy issued in the steady state
would wait on the copy engine behind ~190 MB member batches.  The
persistent column kernel leaves a few SMs free (`sm_reserve`) so the bandwidth-bound transposes of
the other two stages are not queued behind it.  A slab needs (i) one read-only halo row above it
for the 4-point IDW stencil of H, uploaded with it, and (ii) the Y' rows of observations within
`radius` rows of its edges, which its neighbours pack right after their own H(x)
(`mdc_obs_pack_rows`) and it appends (`mdc_obs_append_rows`) -- the same halo mechanism as the
multi-GPU sharding in parallel.py, here between slabs on one device.  Results are bit-identical to
the one-shot analysis (candidates are ordered by (global cell, global observation id)).

Host-side plumbing only; every byte of arithmetic happens in the C-ABI library.
"""
from __future__ import annotations

import copy
import math
import queue
import threading
import time

import numpy as np

from . import capi


class StreamedLetkf:
    def __init__(self, device, gnx, gny, nz, k, radius, slab_rows=32, slots=4, sm_reserve=8, workers=None,
                 row_range=None):
        """row_range = (Y0, Y1): analyse only the global rows [Y0, Y1) (one rank's share of a
        column-sharded job, parallel.py); default: the whole grid."""
        self.gnx, self.gny, self.nz, self.k = gnx, gny, nz, k
        self.reach = int(math.floor(radius))
        self.Y0, self.Y1 = (0, gny) if row_range is None else row_range
        rows = self.Y1 - self.Y0
        self.nslab = max(1, (rows + slab_rows - 1) // slab_rows)
        self.bounds = [(self.Y0 + (rows * s) // self.nslab, self.Y0 + (rows * (s + 1)) // self.nslab)
                       for s in range(self.nslab)]
        # slabs whose observations a slab's columns can reach: every slab within `reach` rows, not only the two
        # adjacent ones (a slab lower than the reach leaves the slab after next inside it).  self.depth = how many
        # slabs ahead must have gone through H before a slab is analysed; they all hold a slot meanwhile.
        self.neigh = [[d for d, (d0, d1) in enumerate(self.bounds)
                       if d != s and d0 < y1 + self.reach and d1 > y0 - self.reach]
                      for s, (y0, y1) in enumerate(self.bounds)]
        self.depth = max([abs(d - s) for s, ds in enumerate(self.neigh) for d in ds] or [0])
        self.nslots = max(3, slots if workers is None else workers, self.depth + 2) if self.nslab > 1 else 1
        self.sm_reserve = sm_reserve
        self.ctxs = [capi.Context(device) for _ in range(self.nslots)]
        self._ens = [None] * self.nslots
        self._obs = [None] * self.nslab         # one observation store per slab, on the context of slot s % nslots
        self._pool = None
        self.trace = []

    def close(self):
        for w in range(len(self.ctxs)):
            if self._ens[w] is not None:
                self._ens[w].close()
        for ob in self._obs:
            if ob is not None:
                ob.close()
        if self._pool:
            self.ctxs[0].dev_free(self._pool[0])
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def analyse(self, member_ptrs, obs, params, host_row0=0, host_ny=None, ext_top=(), ext_bottom=()):
        """member_ptrs: k host pointers (pinned for full PCIe speed) to [nz][host_ny][gnx] float64
        arrays whose first row is global row host_row0 (default: the whole grid), updated IN PLACE.
        obs: dict of global observation arrays.  ext_top / ext_bottom: (device pointer, rows) of packed
        observation rows received from the ranks above / below this row range; they are appended to
        the slabs within reach of that edge.  Returns summed stats."""
        import metada_b200 as mb
        S, R, K, D = self.nslab, self.reach, self.nslots, self.depth
        if int(math.floor(params.radius)) > R:
            raise ValueError(f"StreamedLetkf was planned for radius < {R + 1}; analyse() got {params.radius}: the "
                             "observation halos would miss rows")
        host_ny = self.gny if host_ny is None else host_ny
        prm = copy.copy(params)
        if S > 1:
            prm.sm_reserve = self.sm_reserve
        gid_all = np.arange(len(obs["y"]), dtype=np.int64)
        own_idx, halo_n = [], []
        for s, (y0, y1) in enumerate(self.bounds):
            lo = y0 if y0 > 0 else -(1 << 30)
            hi = y1 if y1 < self.gny else (1 << 30)
            idx = np.nonzero((obs["y"] >= lo) & (obs["y"] < hi))[0]
            own_idx.append(idx)
            ys = obs["y"][idx]
            cnt = {}
            for dst in self.neigh[s]:
                d0, d1 = self.bounds[dst]
                cnt[dst] = (d0 - R, d1 + R, int(np.count_nonzero((ys >= d0 - R) & (ys < d1 + R))))
            halo_n.append(cnt)
        rd = self.k + 8
        max_rows = max(y1 - y0 for (y0, y1) in self.bounds) + 1
        need = sum(n for c in halo_n for (_, _, n) in c.values()) * rd * 8
        if not self._pool or self._pool[1] < need:
            if self._pool:
                self.ctxs[0].dev_free(self._pool[0])
            self._pool = (self.ctxs[0].dev_malloc(max(need, 8)), need)
        for w in range(K):
            if self._ens[w] is None:
                self._ens[w] = mb.Ensemble(self.ctxs[w], self.gnx, max_rows, self.nz, self.k)
        # Every slab's observations go to the device now, before the member traffic starts: a small copy issued later
        # would queue on the copy engine behind ~190 MB member batches (measured: 37 ms per slab instead of < 2 ms).
        # Slots are handed out in FIFO order, so slab s always lands in slot s % K.
        for s in range(S):
            idx = own_idx[s]
            args = (obs["x"][idx], obs["y"][idx], obs["z"][idx], obs["value"][idx], obs["err"][idx], obs["valid"][idx])
            if self._obs[s] is None:
                self._obs[s] = mb.Observations(self.ctxs[s % K], *args, gid=gid_all[idx])
            else:
                self._obs[s].assign(*args, gid=gid_all[idx])
        halo_buf = [dict() for _ in range(S)]          # slab s -> {dst: (devptr, nrows)}
        pool_off = [0]
        stats = [None] * S
        errors = []
        free_slots, q_up, q_down = queue.Queue(), queue.Queue(), queue.Queue()
        for w in range(K):
            free_slots.put(w)
        self.trace = []
        t_base = time.perf_counter()

        def fail(e):
            errors.append(e)
            q_up.put(None); q_down.put(None); free_slots.put(None)

        def uploader():
            try:
                for s in range(S):
                    w = free_slots.get()
                    if w is None or errors:
                        return
                    if w != s % K:
                        raise RuntimeError(f"pipeline slots out of order: slab {s} got slot {w}")
                    t0 = time.perf_counter()
                    y0, y1 = self.bounds[s]
                    ens = self._ens[w]
                    ens.set_rows((y1 - y0) + (1 if y1 < self.gny else 0))
                    ens.set_domain(0, y0, self.gnx, self.gny, self.gnx, y1 - y0)
                    ens.upload_rows(member_ptrs, host_ny, y0 - host_row0)
                    self.ctxs[w].sync()
                    self.trace.append(("up", s, w, t0 - t_base, time.perf_counter() - t_base))
                    q_up.put((s, w))
                q_up.put(None)
            except BaseException as e:  # noqa: BLE001
                fail(e)

        slot_of = {}

        def do_hx(s):
            w = slot_of[s]
            ens = self._ens[w]
            ob = self._obs[s]
            ob.hx(ens)
            for dst, (lo, hi, n) in halo_n[s].items():
                if n > 0:
                    p = self._pool[0] + pool_off[0]
                    pool_off[0] += n * rd * 8
                    got = ob.pack_rows(lo, hi, p, n)
                    assert got == n, (got, n)
                    halo_buf[s][dst] = (p, n)

        def compute():
            try:
                loaded = -1
                for s in range(S):
                    while loaded < min(s + D, S - 1):      # slabs s+1 .. s+D must be on the device: their H feeds s
                        item = q_up.get()
                        if item is None or errors:
                            return
                        slot_of[item[0]] = item[1]
                        loaded = item[0]
                        t0 = time.perf_counter()
                        do_hx(loaded)
                        self.trace.append(("hx", loaded, item[1], t0 - t_base, time.perf_counter() - t_base))
                    w = slot_of[s]
                    ob = self._obs[s]
                    t0 = time.perf_counter()
                    for src in self.neigh[s]:
                        if s in halo_buf[src]:
                            p, n = halo_buf[src][s]
                            ob.append_rows(p, n)
                    if self.bounds[s][0] - R < self.Y0:        # rows from other ranks: supersets are harmless
                        for p, n in ext_top:
                            ob.append_rows(p, n)
                    if self.bounds[s][1] + R > self.Y1:
                        for p, n in ext_bottom:
                            ob.append_rows(p, n)
                    stats[s] = capi.letkf_analyse(self._ens[w], ob, prm)
                    self.trace.append(("an", s, w, t0 - t_base, time.perf_counter() - t_base))
                    q_down.put((s, w))
                q_down.put(None)
            except BaseException as e:  # noqa: BLE001
                fail(e)

        def downloader():
            try:
                while True:
                    item = q_down.get()
                    if item is None or errors:
                        return
                    s, w = item
                    t0 = time.perf_counter()
                    y0, y1 = self.bounds[s]
                    self._ens[w].download_rows(member_ptrs, host_ny, y0 - host_row0, y1 - y0)
                    self.trace.append(("dn", s, w, t0 - t_base, time.perf_counter() - t_base))
                    free_slots.put(w)
            except BaseException as e:  # noqa: BLE001
                fail(e)

        threads = [threading.Thread(target=f) for f in (uploader, compute, downloader)]
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
