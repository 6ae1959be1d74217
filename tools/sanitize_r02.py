"""Round-2 code paths for compute-sanitizer (memcheck / racecheck / initcheck), small sizes: the restructured packed
kernel (bulk copies, w = Z (Z g) through shared partials), accurate observations (long schedules, kappa_max), the
per-level passes (first pass, second observation-space pass for 24 < p <= 32, work list), REF modes on geographic
observations, and the sharded geographic analysis (geography windows, box packing, extended halo rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from metada_b200.parallel import GeoSlabLetkf

ctx = mb.Context(0)


def grid_case(nx, ny, nz, k, P, radius, rv=0.0, sigma=0.1, **kw):
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(syn.ensemble(k, nx, ny, nz, seed=k))
    o = syn.observations(P, nx, ny, nz, seed=k + 1, sigma=sigma)
    o["err"][:] = sigma
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(radius, 1.02, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv, **kw))
    assert np.isfinite(ens.download()).all() and st["numeric_failures"] == 0, st
    print("grid", k, rv, sigma, kw, st["columns"], st["max_sweeps"], st["small_transforms"], st["redo_transforms"], flush=True)
    ens.close(); obs.close()


grid_case(12, 10, 3, 80, 130, 5.0)                       # packed kernel, nt = 10 (compile-time products)
grid_case(10, 9, 2, 40, 100, 5.0)                        # four CTAs per SM
grid_case(9, 8, 2, 104, 90, 5.0)                         # generic tile walk
grid_case(10, 9, 2, 80, 110, 5.0, sigma=0.01)            # long schedule (condition bound ~ 5e3)
grid_case(10, 9, 2, 80, 110, 5.0, sigma=0.01, kappa_max=500.0)   # ... sent to the redo list instead
grid_case(12, 11, 6, 64, 1500, 6.0, rv=2.0)              # per-level: first pass, second pass (24 < p <= 32), work list
grid_case(10, 9, 4, 128, 700, 6.0, rv=1.5)

vc = np.array([1000.0, 850.0])
for mode, loc in ((mb.MODE_REF_COMPAT, mb.LOC_CUTOFF), (mb.MODE_REF_ETKF, mb.LOC_CUTOFF)):
    nx, ny, nz, k, P = 14, 11, 2, 12, 160
    lat, lon = syn.geography(nx, ny, lon0=178.9)
    o = syn.geo_observations(P, lat, lon, vc, seed=5, margin=0.4)
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(syn.ensemble(k, nx, ny, nz, seed=6))
    ens.set_geography(lat, lon, vc)
    obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(60.0, 1.1, mode, loc))
    assert np.isfinite(ens.download()).all()
    print("geo ref mode", mode, st["columns"], st["max_local_obs"], flush=True)
    ens.close(); obs.close()

# sharded geographic analysis, three ranks played in this process
nx, ny, nz, k, P, world = 16, 15, 4, 24, 300, 3
lat, lon = syn.geography(nx, ny)
vc3 = np.array([1000.0, 850.0, 500.0])
o = syn.geo_observations(P, lat, lon, vc3, seed=9)
o = dict(o)
o["var"] = np.random.default_rng(3).integers(0, 2, P).astype(np.int32)
X = syn.ensemble(k, nx, ny, nz, seed=10)
jobs = [GeoSlabLetkf(ctx, lat, lon, vc3, nz, k, r, world, 50.0, [3, 1]) for r in range(world)]
sends = []
for job in jobs:
    job.ens.upload(np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :]))
    job.set_observations(o)
    sends.append(job.pack_halo())
params = capi.make_params(50.0, 1.02, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
for r, job in enumerate(jobs):
    st = job.analyse(params, recv={src: sends[src][r] for src in range(world) if src != r})
    assert np.isfinite(job.ens.download()).all()
    print("geo shard", r, st["columns"], job.halo_rows_last, flush=True)
for job in jobs:
    job.close()
torch.cuda.synchronize()
ctx.close()
print("ok")
