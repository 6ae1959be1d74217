"""Timings of the WRF-shaped case on the GPU box: geographic observations (haversine selection, nearest-grid-point
location) on a multi-variable state.  Writes gpurun_out/geo_probe.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn


def main():
    ctx = mb.Context(0)
    out = []
    for name, nx, ny, var_nlev, k, P, radius, dlat, dlon in (("wrf-300", 300, 300, [30, 30, 30, 1], 40, 60000, 40.0, 0.09, 0.11),
                                                             ("wrf-600", 600, 600, [30, 30, 1], 80, 250000, 20.0, 0.045, 0.055)):
        nz = sum(var_nlev)
        lat, lon = syn.geography(nx, ny, dlat=dlat, dlon=dlon)
        vc = np.linspace(1000.0, 100.0, max(var_nlev))
        o = syn.geo_observations(P, lat, lon, vc, seed=42)
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        ens.set_geography(lat, lon, vc)
        ens.set_variables(var_nlev)
        obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
        obs.set_variables(np.random.default_rng(1).integers(0, len(var_nlev), P).astype(np.int32))
        rec = {"case": name, "nx": nx, "ny": ny, "var_nlev": var_nlev, "k": k, "P": P, "radius_km": radius, "grid_spacing_deg": [dlat, dlon]}
        for it in range(3):
            ens.fill_synthetic(1000)
            ctx.sync()
            ctx.timer_start(); obs.locate(ens); rec["ms_locate"] = ctx.timer_stop()
            st = capi.letkf_analyse(ens, obs, capi.make_params(radius, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
        rec.update({key: st[key] for key in ("ms_hx", "ms_index", "ms_columns", "ms_total", "columns", "max_local_obs",
                                             "small_transforms", "redo_transforms")})
        rec["mean_local_obs"] = st["sum_local_obs"] / st["columns"]
        rec["columns_per_s"] = st["columns"] / (st["ms_total"] * 1e-3)
        # the same state and observation density with GRID coordinates (integer distances) for comparison
        og = syn.observations(P, nx, ny, max(var_nlev), seed=42)
        gobs = mb.Observations(ctx, og["x"], og["y"], og["z"], og["value"], og["err"], og["valid"])
        for it in range(2):
            ens.fill_synthetic(1000)
            sg = capi.letkf_analyse(ens, gobs, capi.make_params(4.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
        rec["grid_coordinates_same_density"] = {"ms_columns": sg["ms_columns"], "mean_local_obs": sg["sum_local_obs"] / sg["columns"]}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        obs.close(); gobs.close(); ens.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/geo_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
