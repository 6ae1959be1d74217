"""First-contact probe on the GPU box: host info, FP64/HBM microbenchmarks, quick LETKF timings."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn


def sh(cmd):
    try:
        return subprocess.check_output(cmd, shell=True, text=True, stderr=subprocess.STDOUT).strip()
    except Exception as e:  # noqa: BLE001
        return f"ERR {e}"


def main():
    out = {"nproc": os.cpu_count(), "free": sh("free -g | head -2"),
           "smi": sh("nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.limit --format=csv")}
    ctx = mb.Context(0)
    out["sm_count"] = ctx.sm_count()
    out["fp64_fma_tflops"] = ctx.bench_fp64_fma()
    out["fp64_dmma_tflops"] = ctx.bench_fp64_dmma()
    out["hbm_copy_gbs"] = ctx.bench_hbm_copy()
    print(json.dumps(out, indent=1), flush=True)
    cases = [("C1", 100, 100, 1, 20, 1000, 10.0), ("C3q", 200, 200, 50, 40, 25000, 7.0),
             ("C5q", 256, 256, 60, 80, 29100, 8.0), ("C4q", 128, 128, 20, 128, 7300, 8.0)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases.append(("C3", 400, 400, 50, 40, 100000, 7.0))
    for name, nx, ny, nz, k, P, r in cases:
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        o = syn.observations(P, nx, ny, nz, seed=42)
        for mode, mname, solver in ((mb.MODE_CANONICAL, "canonical-ns", 0), (mb.MODE_CANONICAL, "canonical-ns-full", 3),
                                    (mb.MODE_CANONICAL, "canonical-jacobi", 1),
                                    (mb.MODE_REF_ETKF, "ref_etkf", 0), (mb.MODE_REF_COMPAT, "ref_compat", 0)):
            if solver == 3 and not 24 <= k <= 80:
                continue
            reps = 3 if mode == mb.MODE_CANONICAL else 1
            best = None
            for _ in range(reps):
                ens.fill_synthetic(1000)
                obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
                p = capi.make_params(r, 1.0, mode, mb.LOC_GASPARI_COHN, solver=solver)
                t0 = time.time()
                st_ = capi.letkf_analyse(ens, obs, p)
                wall_ = time.time() - t0
                if best is None or st_["ms_columns"] < best[0]["ms_columns"]:
                    best = (st_, wall_)
                if _ < reps - 1:
                    obs.close()
            st, wall = best
            cols = st["columns"]
            print(json.dumps({"case": name, "mode": mname, "cols": cols, "wall_s": round(wall, 4),
                              "ms_hx": st["ms_hx"], "ms_index": st["ms_index"], "ms_columns": st["ms_columns"],
                              "cols_per_s": cols / (st["ms_total"] * 1e-3), "mean_ploc": st["sum_local_obs"] / cols,
                              "max_ploc": st["max_local_obs"], "mean_sweeps": st["sum_sweeps"] / cols,
                              "max_sweeps": st["max_sweeps"], "redo": st["redo_transforms"]}), flush=True)
            obs.close()
        ens.close()
    ctx.close()


if __name__ == "__main__":
    main()
