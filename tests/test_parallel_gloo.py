"""CPU, world_size > 1 over gloo: the column-sharding plumbing (metada_b200/parallel.py) -- slab
ownership, halo plan, row exchange -- with the oracle standing in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metada_b200 import parallel as par
from metada_b200 import synthetic as syn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_slab_ownership_partitions_every_row():
    for gny in (7, 24, 1500):
        for world in (1, 2, 3, 4, 8):
            if world > gny:
                continue
            y = np.arange(-5, gny + 5)
            own = par.owner_of_row(y, gny, world)
            assert own.min() == 0 and own.max() == world - 1
            for r in range(world):
                y0, y1 = par.slab_bounds(gny, r, world)
                inside = np.clip(y, 0, gny - 1)
                assert np.array_equal(own == r, (inside >= y0) & (inside < y1))


@pytest.mark.parametrize("gny,world,reach", [(24, 2, 4), (24, 3, 5), (10, 4, 6), (1500, 8, 8)])
def test_halo_plan_covers_every_needed_row(gny, world, reach):
    plan = par.halo_plan(gny, world, reach)
    for dst in range(world):
        d0, d1 = par.slab_bounds(gny, dst, world)
        for y in range(-3, gny + 3):
            src = int(par.owner_of_row(y, gny, world))
            needed = (d0 - reach <= y < d1 + reach) and src != dst
            lo, hi = plan.get((src, dst), (0, 0))
            assert (lo <= y < hi) == needed, (dst, y, src, lo, hi)


def _worker(rank, world, port, nx, ny, nz, k, P, radius, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    X = syn.ensemble(k, nx, ny, nz, seed=3)
    o = syn.observations(P, nx, ny, nz, seed=4, sigma=0.2)
    o["y"][:3] = [-2, ny + 1, ny - 1]              # out-of-grid obs belong to the edge slabs
    own = par.select_own(o, ny, rank, world)
    # the rank's own Y' rows (H on its own observations only)
    _, _, Yp, d = orc.obs_space(X, o["x"][own], o["y"][own], o["z"][own], o["value"][own])
    rd = k + 8
    rows = np.concatenate([Yp, d[:, None], o["value"][own, None], o["err"][own, None], np.ones((len(own), 1)),
                           o["x"][own, None], o["y"][own, None], o["z"][own, None], own[:, None]], axis=1)
    plan = par.halo_plan(ny, world, int(np.floor(radius)))
    send = {}
    for (src, dst), (lo, hi) in plan.items():
        if src == rank:
            m = (o["y"][own] >= lo) & (o["y"][own] < hi)
            send[dst] = torch.from_numpy(np.ascontiguousarray(rows[m]))
    recv = par.exchange_rows(dist, rank, world, send, rd, torch.device("cpu"))
    allrows = np.concatenate([rows] + [t.numpy() for t in recv]) if recv else rows
    gid = allrows[:, k + 7].astype(np.int64)
    assert len(np.unique(gid)) == len(gid)         # no observation delivered twice
    # received rows must be the sender's bits
    Yp_all = orc.obs_space(X, o["x"], o["y"], o["z"], o["value"])[2]
    assert np.array_equal(allrows[:, :k], Yp_all[gid])
    # analyse the rank's own columns with own + halo observations only
    y0, y1 = par.slab_bounds(ny, rank, world)
    cols = np.array([y * nx + x for y in range(y0, y1) for x in range(nx)], np.int64)
    order = np.argsort(gid)
    loc = orc.letkf(X, allrows[order, k + 4].astype(np.int32), allrows[order, k + 5].astype(np.int32),
                    allrows[order, k + 6].astype(np.int32), allrows[order, k + 1], allrows[order, k + 2],
                    radius=radius, cols=cols)
    out[rank] = loc["Xa"][:, :, y0:y1, :].copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_analysis_equals_single_process(world):
    from oracle import orc
    nx, ny, nz, k, P, radius = 18, 21, 2, 6, 90, 4.0
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, nz, k, P, radius, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X = syn.ensemble(k, nx, ny, nz, seed=3)
    o = syn.observations(P, nx, ny, nz, seed=4, sigma=0.2)
    o["y"][:3] = [-2, ny + 1, ny - 1]
    full = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=radius)["Xa"]
    for r in range(world):
        y0, y1 = par.slab_bounds(ny, r, world)
        # same observation SETS in the same (global id) order -> bit-identical
        assert np.array_equal(out[r], full[:, :, y0:y1, :]), r


def test_geographic_halo_boxes_cover_every_reachable_observation():
    """The boxes GeoSlabLetkf sends by (bounding box of a slab's columns in the geography's frame, widened by the
    reach of the radius) contain every observation within the radius of any of the slab's columns -- checked against
    the oracle's haversine (Location.hpp:349-357) on a curvilinear grid across the date line."""
    from oracle import orc
    lat, lon = syn.geography(40, 33, lon0=178.2, lat0=-44.0)
    o = syn.geo_observations(500, lat, lon, None, seed=12)
    lon_c = float(np.degrees(np.arctan2(np.sin(np.radians(lon)).sum(), np.cos(np.radians(lon)).sum())))
    u = (lon - lon_c) - 360.0 * np.rint((lon - lon_c) / 360.0)
    frame = {"lon_c": lon_c, "umin": float(u.min()), "umax": float(u.max()), "latmin": float(lat.min()), "latmax": float(lat.max())}
    radius, world = 70.0, 3
    boxes = par.geo_halo_boxes(lat, lon, frame, world, radius)
    ou = (o["lon"] - lon_c) - 360.0 * np.rint((o["lon"] - lon_c) / 360.0)
    reached = 0
    for r, (la0, la1, u0, u1) in enumerate(boxes):
        y0, y1 = par.slab_bounds(lat.shape[0], r, world)
        inbox = (o["lat"] >= la0) & (o["lat"] <= la1) & (ou >= u0) & (ou <= u1)
        for y in range(y0, y1, 3):
            for x in range(0, lat.shape[1], 4):
                sel = orc.select_local_geo(lat[y, x], lon[y, x], o["lat"], o["lon"], radius)
                assert inbox[sel].all(), (r, y, x)
                reached += len(sel)
        assert inbox.sum() < len(inbox)          # ... and it is a proper subset
    assert reached > 1000


def _geo_worker(rank, world, port, nx, ny, nz, k, P, radius, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    lat, lon = syn.geography(nx, ny, lon0=178.4)
    vc = np.array([1000.0, 850.0])
    o = syn.geo_observations(P, lat, lon, vc, seed=14)
    X = syn.ensemble(k, nx, ny, nz, seed=15)
    ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    lon_c = float(np.degrees(np.arctan2(np.sin(np.radians(lon)).sum(), np.cos(np.radians(lon)).sum())))
    u = (lon - lon_c) - 360.0 * np.rint((lon - lon_c) / 360.0)
    frame = {"lon_c": lon_c, "umin": float(u.min()), "umax": float(u.max()), "latmin": float(lat.min()), "latmax": float(lat.max())}
    boxes = par.geo_halo_boxes(lat, lon, frame, world, radius)
    own = np.nonzero(par.owner_of_row(ey, ny, world) == rank)[0]        # ownership = the slab of the LOCATED row
    ou = (o["lon"] - lon_c) - 360.0 * np.rint((o["lon"] - lon_c) / 360.0)
    # what travels: the global id (the rows' other fields are functions of it here; the device rows carry them)
    send = {}
    for dst in range(world):
        if dst != rank:
            la0, la1, u0, u1 = boxes[dst]
            m = (o["lat"][own] >= la0) & (o["lat"][own] <= la1) & (ou[own] >= u0) & (ou[own] <= u1)
            send[dst] = torch.from_numpy(own[m].astype(np.float64)[:, None].copy())
    recv = par.exchange_rows(dist, rank, world, send, 1, torch.device("cpu"))
    have = np.sort(np.concatenate([own] + [t.numpy()[:, 0].astype(np.int64) for t in recv]))
    assert len(np.unique(have)) == len(have)
    # analyse with own + received observations only (ascending global id = the order of the one-store run)
    r = orc.letkf_ext(X, ex[have], ey[have], ez[have], o["value"][have], o["err"][have], o["valid"][have], radius=radius,
                      glat=lat, glon=lon, olat=o["lat"][have], olon=o["lon"][have])
    y0, y1 = par.slab_bounds(ny, rank, world)
    out[rank] = (r["Xa"][:, :, y0:y1, :].copy(), len(own), len(have))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_geographic_plan_equals_single_process(world):
    """The plan of GeoSlabLetkf / mdc_geo_sharded_analyse over gloo with the oracle as the kernel: ownership by the
    located row, halo observations by box.  Every rank's rows equal the single-process analysis bit for bit although
    it only ever sees its own and the received observations."""
    from oracle import orc
    nx, ny, nz, k, P, radius = 22, 21, 2, 6, 260, 55.0
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    procs = [ctx.Process(target=_geo_worker, args=(r, world, port, nx, ny, nz, k, P, radius, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    lat, lon = syn.geography(nx, ny, lon0=178.4)
    vc = np.array([1000.0, 850.0])
    o = syn.geo_observations(P, lat, lon, vc, seed=14)
    X = syn.ensemble(k, nx, ny, nz, seed=15)
    ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    full = orc.letkf_ext(X, ex, ey, ez, o["value"], o["err"], o["valid"], radius=radius, glat=lat, glon=lon,
                         olat=o["lat"], olon=o["lon"])["Xa"]
    assert sum(out[r][1] for r in range(world)) == P
    for r in range(world):
        y0, y1 = par.slab_bounds(ny, r, world)
        assert np.array_equal(out[r][0], full[:, :, y0:y1, :]), r
        assert out[r][2] < P                                  # a rank does not need every observation
