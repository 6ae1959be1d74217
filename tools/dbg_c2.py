"""C2 (global EnKF) wall time per call in a fresh process, optionally after a per-level LETKF."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
ctx = mb.Context(0)
if len(sys.argv) > 1:
    e2 = mb.Ensemble(ctx, 48, 48, 60, 128)
    o2 = syn.observations(1152, 48, 48, 60, seed=42)
    ob2 = mb.Observations(ctx, o2["x"], o2["y"], o2["z"], o2["value"], o2["err"], o2["valid"])
    e2.fill_synthetic(1000)
    capi.letkf_analyse(e2, ob2, capi.make_params(8.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=float(sys.argv[1])))
    ob2.close(); e2.close()
nx, ny, k, P = 400, 250, 40, 10000
ens = mb.Ensemble(ctx, nx, ny, 1, k)
o = syn.observations(P, nx, ny, 1, seed=42, distinct=True)
Z = np.random.default_rng(7).standard_normal((P, k))
ts = []
for _ in range(6):
    ens.fill_synthetic(1000)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    ctx.sync(); t0 = time.perf_counter()
    capi.enkf_analyse(ens, obs, 1.0, Z=Z, want_gain_stats=False)
    ctx.sync(); ts.append(round(1e3 * (time.perf_counter() - t0), 2))
    obs.close()
print("enkf ms per call:", ts)
