import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
nx, ny, nz, k, P, r = 256, 256, 60, 80, 29100, 8.0
ctx = mb.Context(0)
ens = mb.Ensemble(ctx, nx, ny, nz, k)
o = syn.observations(P, nx, ny, nz, seed=42)
for rep in range(2):
    ens.fill_synthetic(1000)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(r, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
    raw = ctx.last_stats()
    obs.close()
cols = st["columns"]
print(json.dumps({"ms": st["ms_columns"], "products_per_col": st["sum_sweeps"] / cols,
                  "cycles_in_products_per_column_by_warp": [raw[8 + w] / cols for w in range(8)]}))
