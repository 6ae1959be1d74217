"""Small geographic / multi-variable runs for compute-sanitizer (memcheck / racecheck): locate (ring walk and scan),
lattice index, EXT column kernels (Jacobi, packed Newton-Schulz, observation space), row-slab transfers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn

ctx = mb.Context(0)
vc = np.array([1000.0, 850.0, 500.0])
for k, var_nlev, P, radius, rv in ((12, [3], 150, 60.0, 0.0), (40, [3, 3, 1], 200, 55.0, 0.0), (64, [3], 40, 30.0, 0.0),
                                   (32, [3, 1], 150, 60.0, 1.5), (80, [3], 200, 70.0, 0.0)):
    nx, ny, nz = 14, 11, sum(var_nlev)
    lat, lon = syn.geography(nx, ny, lon0=178.9)
    o = syn.geo_observations(P, lat, lon, vc, seed=k, margin=0.4)
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(syn.ensemble(k, nx, ny, nz, seed=k))
    ens.set_geography(lat, lon, vc)
    ens.set_variables(var_nlev)
    obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
    obs.set_variables(np.random.default_rng(k).integers(0, len(var_nlev), P).astype(np.int32))
    os.environ["MDC_GEO_LOCATE_BRUTE"] = "1"
    obs.locate(ens)
    a = obs.grid_coords()
    del os.environ["MDC_GEO_LOCATE_BRUTE"]
    obs.locate(ens)
    b = obs.grid_coords()
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    counts = obs.query_counts(ens, radius)
    st = capi.letkf_analyse(ens, obs, capi.make_params(radius, 1.02, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv))
    assert np.isfinite(ens.download()).all()
    print(k, var_nlev, rv, st["columns"], int(counts.max()), st["small_transforms"], st["numeric_failures"], flush=True)
    ens.close(); obs.close()
# row-slab transfers (double-buffered staging)
nx, ny, nz, k = 20, 12, 3, 20
X = syn.ensemble(k, nx, ny, nz, seed=3)
ens = mb.Ensemble(ctx, nx, 5, nz, k)
ens.set_domain(0, 4, nx, ny, nx, 5)
ens.upload_rows([X[m].ctypes.data for m in range(k)], ny, 4)
ctx.sync()
out = np.zeros_like(X)
ens.download_rows([out[m].ctypes.data for m in range(k)], ny, 4, 5)
assert np.array_equal(out[:, :, 4:9], X[:, :, 4:9])
ens.close(); ctx.close()
print("ok")
