"""WRF-shaped case with vertical localisation (one transform per level of every variable): geographic observations,
multi-variable state, radius_v > 0 (development probe for the per-level passes on EXT input)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
ctx = mb.Context(0)
nx, ny, var_nlev, k, P, radius, rv = 200, 200, [30, 30, 1], 40, 40000, 40.0, 3.0
nz = sum(var_nlev)
lat, lon = syn.geography(nx, ny, dlat=0.09, dlon=0.11)
vc = np.linspace(1000.0, 100.0, max(var_nlev))
o = syn.geo_observations(P, lat, lon, vc, seed=42)
ens = mb.Ensemble(ctx, nx, ny, nz, k)
ens.set_geography(lat, lon, vc)
ens.set_variables(var_nlev)
obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
obs.set_variables(np.random.default_rng(1).integers(0, len(var_nlev), P).astype(np.int32))
for it in range(3):
    ens.fill_synthetic(1000)
    st = capi.letkf_analyse(ens, obs, capi.make_params(radius, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv))
s, s2 = ens.checksum()
print(json.dumps({"case": f"{nx}x{ny} {var_nlev} k={k} P={P} r={radius} km r_v={rv}", "ms_columns": st["ms_columns"], "ms_total": st["ms_total"],
                  "transforms": nx * ny * nz, "small": st["small_transforms"], "redo": st["redo_transforms"], "fail": st["numeric_failures"],
                  "checksum": [s, s2]}))
