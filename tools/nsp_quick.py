"""Quick timing of the packed column kernel on the probe shapes (best of 3, device-timed ms of the column phase)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn

CASES = {"C5q": (256, 256, 60, 80, 29100, 8.0, 0.0), "C3q": (200, 200, 50, 40, 25000, 7.0, 0.0),
         "C4q": (128, 128, 60, 128, 8192, 8.0, 0.0), "C4v": (96, 96, 60, 128, 4608, 8.0, 5.0),
         "C1": (100, 100, 1, 20, 1000, 10.0, 0.0)}


def main():
    names = sys.argv[1:] or ["C5q", "C3q", "C4q"]
    ctx = mb.Context(0)
    for name in names:
        nx, ny, nz, k, P, r, rv = CASES[name]
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        o = syn.observations(P, nx, ny, nz, seed=42)
        best = None
        for _ in range(3):
            ens.fill_synthetic(1000)
            obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
            st = capi.letkf_analyse(ens, obs, capi.make_params(r, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv))
            obs.close()
            if best is None or st["ms_columns"] < best["ms_columns"]:
                best = st
        cols = best["columns"]
        print(json.dumps({"case": name, "ms_columns": round(best["ms_columns"], 3), "Mcols_per_s": round(cols / best["ms_columns"] / 1e3, 4),
                          "mean_ploc": round(best["sum_local_obs"] / cols, 2), "mean_products": round(best["sum_sweeps"] / cols, 3),
                          "redo": best["redo_transforms"], "small": best["small_transforms"], "fail": best["numeric_failures"]}), flush=True)
        ens.close()
    ctx.close()


if __name__ == "__main__":
    main()
