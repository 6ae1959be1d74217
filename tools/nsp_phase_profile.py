"""Per-phase time of the packed Newton-Schulz column kernel from its own clock64() ticks.
Build the nt = 10 unit with the ticks first:
    MDC_NSP_PROFILE=1 MDC_BUILD_ONLY=nsp_10_10.o python -c "import metada_b200 as mb; mb.build_library()"
then run this on the GPU box; rebuild without MDC_NSP_PROFILE afterwards (the ticks cost a few percent)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn

if os.environ.get("NSP_PROFILE_UPDATE"):
    NAMES = {8: "selection..start", 9: "iteration", 10: "update products", 11: "state wait", 12: "level means", 13: "w=ZZg",
             14: "stores+mean", 15: "column total"}
else:
  NAMES = {8: "selection", 9: "gather+syrk", 10: "norm/A/start", 11: "products", 12: "iteration epilogues", 13: "w=ZZg",
         14: "update", 15: "column total"}


def main():
    nx, ny, nz, k, P, r = 256, 256, 60, 80, 29100, 8.0
    if len(sys.argv) > 1:
        nx, ny, nz, k, P, r = [float(v) if i == 5 else int(v) for i, v in enumerate(sys.argv[1:7])]
    ctx = mb.Context(0)
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    o = syn.observations(P, nx, ny, nz, seed=42)
    out = {}
    for rep in range(2):
        ens.fill_synthetic(1000)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        st = capi.letkf_analyse(ens, obs, capi.make_params(r, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
        raw = ctx.last_stats()
        obs.close()
    cols = st["columns"]
    out = {"case": f"{nx}x{ny}x{nz} k={k} P={P} r={r}", "ms_columns": st["ms_columns"], "columns": cols,
           "mean_products": st["sum_sweeps"] / cols, "small": st["small_transforms"],
           "ticks_per_column": {NAMES[i]: raw[i] / max(1, cols - st["small_transforms"]) for i in range(8, 16)}}
    tot = raw[15]
    out["share"] = {NAMES[i]: raw[i] / tot for i in range(8, 15)} if tot else None
    print(json.dumps(out, indent=1))
    ens.close(); ctx.close()


if __name__ == "__main__":
    main()
