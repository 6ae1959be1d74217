"""Streamed host-buffer pipeline (the bench's e2e leg) on the full C5 workload under different settings:
slots, slab height, SMs reserved for the member transposes.  Writes gpurun_out/e2e_probe.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn


def main():
    nx, ny, nz, k, P, radius = 1500, 1500, 60, 80, 1000000, 8.0
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        ny, P = 320, 213000
    ctx = mb.Context(0)
    n = nx * ny * nz
    t0 = time.perf_counter()
    host = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(k)]
    ptrs = [t.data_ptr() for t in host]
    out = {"pin_s": time.perf_counter() - t0, "GB_each_way": n * k * 8 / 1e9, "runs": []}
    o = syn.observations(P, nx, ny, nz, seed=42)
    params = capi.make_params(radius, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
    ens = mb.Ensemble(ctx, nx, ny, nz, k)

    def refill():
        ens.fill_synthetic(1000)
        ens.download_ptrs(0, ptrs)
        ctx.sync()

    configs = [(32, 4, 8), (32, 4, 4), (32, 5, 12), (48, 4, 8)]
    for slab_rows, slots, smr in configs:
        sl = mb.StreamedLetkf(0, nx, ny, nz, k, radius, slab_rows=slab_rows, slots=slots, sm_reserve=smr)
        rec = {"slab_rows": slab_rows, "slots": slots, "sm_reserve": smr, "s": []}
        for it in range(2):               # every configuration allocates its own slots and stores: first pass = warm-up
            refill()
            t0 = time.perf_counter()
            sl.analyse(ptrs, o, params)
            rec["s"].append(time.perf_counter() - t0)
        dur = {}
        for kind, s, w, a, b in sl.trace:
            dur.setdefault(kind, []).append(b - a)
        rec["stage_busy_s"] = {kk: sum(v) for kk, v in dur.items()}
        rec["stage_mean_ms"] = {kk: 1e3 * sum(v) / len(v) for kk, v in dur.items()}
        rec["columns_per_s"] = nx * ny / rec["s"][-1]
        out["runs"].append(rec)
        print(json.dumps(rec), flush=True)
        sl.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/e2e_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
