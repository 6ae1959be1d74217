"""C4-tile timing with and without the observation-space path, with the solver statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
ctx = mb.Context(0)
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nz, k, rv = 60, 128, 5.0
P = int(0.5 * nx * ny)
ens = mb.Ensemble(ctx, nx, ny, nz, k)
o = syn.observations(P, nx, ny, nz, seed=42)
for rep in range(2):
    ens.fill_synthetic(1000)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(8.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv))
    print({kk: st[kk] for kk in ("ms_columns", "columns", "sum_local_obs", "max_local_obs", "sum_sweeps", "max_sweeps", "redo_transforms", "small_transforms", "numeric_failures")}, flush=True)
    obs.close()
