"""BASELINE.json configs[0..3] on one B200: timings + property checks (no oracle at these sizes)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn

ctx = mb.Context(0)
out = []

def letkf_case(name, nx, ny, nz, k, P, r, rv=0.0, mode=mb.MODE_CANONICAL, solver=0, reps=3):
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    o = syn.observations(P, nx, ny, nz, seed=42)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    best = None
    for _ in range(reps):
        ens.fill_synthetic(1000)
        ctx.sync()
        ctx.timer_start()
        st = capi.letkf_analyse(ens, obs, capi.make_params(r, 1.0, mode, mb.LOC_GASPARI_COHN, radius_v=rv, solver=solver))
        ms = ctx.timer_stop()
        best = ms if best is None else min(best, ms)
    cols = st["columns"]
    out.append({"case": name, "ms": best, "columns_per_s": cols / (best * 1e-3), "mean_ploc": st["sum_local_obs"] / cols,
                "max_ploc": st["max_local_obs"], "mean_iters": st["sum_sweeps"] / cols, "failures": st["numeric_failures"]})
    print(json.dumps(out[-1]), flush=True)
    obs.close(); ens.close()

letkf_case("C1 100x100x1 k=20 P=1e3 r=10 canonical", 100, 100, 1, 20, 1000, 10.0)
letkf_case("C1 ref_compat (LETKF.hpp arithmetic)", 100, 100, 1, 20, 1000, 10.0, mode=mb.MODE_REF_COMPAT)
letkf_case("C3 400x400x50 k=40 P=1e5 r=7 canonical (packed NS on DMMA)", 400, 400, 50, 40, 100000, 7.0)
letkf_case("C3 canonical (Jacobi)", 400, 400, 50, 40, 100000, 7.0, solver=1)
letkf_case("C3 ref_etkf", 400, 400, 50, 40, 100000, 7.0, mode=mb.MODE_REF_ETKF)
letkf_case("C4-tile 96x96x60 k=128 P=4608 r_h=8 r_v=5 canonical (observation-space kernel for p_loc <= 24, packed NS on DMMA otherwise; per-level transforms)", 96, 96, 60, 128, 4608, 8.0, rv=5.0, reps=1)
letkf_case("C4-tile same, horizontal localisation only", 96, 96, 60, 128, 4608, 8.0, reps=2)

# C2: global stochastic EnKF, n = 1e5 (400 x 250), 40 members, 1e4 distinct obs, supplied draws
nx, ny, k, P = 400, 250, 40, 10000
ens = mb.Ensemble(ctx, nx, ny, 1, k)
o = syn.observations(P, nx, ny, 1, seed=42, distinct=True)
Z = np.random.default_rng(7).standard_normal((P, k))
for gain in (False, True):
    best = None
    for _ in range(3):
        ens.fill_synthetic(1000)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        ctx.sync(); t0 = time.perf_counter()
        d = capi.enkf_analyse(ens, obs, 1.0, Z=Z, want_gain_stats=gain)
        ctx.sync(); ms = 1e3 * (time.perf_counter() - t0)
        best = ms if best is None else min(best, ms)
        obs.close()
    out.append({"case": f"C2 EnKF n=1e5 k=40 P=1e4 gain_stats={gain}", "ms_wall_incl_Z_upload": best, "diag": d})
    print(json.dumps(out[-1]), flush=True)
# global ETKF on the same case
best = None
for _ in range(3):
    ens.fill_synthetic(1000)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    ctx.sync(); t0 = time.perf_counter(); capi.etkf_analyse(ens, obs, 1.0); ctx.sync()
    ms = 1e3 * (time.perf_counter() - t0); best = ms if best is None else min(best, ms); obs.close()
out.append({"case": "C2-shape global ETKF", "ms_wall": best}); print(json.dumps(out[-1]), flush=True)
json.dump(out, open("gpurun_out/configs_probe.json", "w"), indent=1)
