import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
nx, ny, nz, k = 1500, 320, 60, 80
P = int(0.4444 * nx * ny)
ctx = mb.Context(0)
ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.fill_synthetic(1000)
n = nx * ny * nz
host = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(k)]
ptrs = [t.data_ptr() for t in host]
ens.download_ptrs(0, ptrs); ctx.sync(); ens.close()
o = syn.observations(P, nx, ny, nz, seed=42)
params = capi.make_params(8.0, 1.0, 2, 1)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sl = mb.StreamedLetkf(0, nx, ny, nz, k, 8.0, slab_rows=32, slots=W)
for it in range(2):
    t0 = time.perf_counter(); sl.analyse(ptrs, o, params); dt = time.perf_counter() - t0
    print("pass", it, "s", dt, "cols/s", nx * ny / dt, "GB each way", n * k * 8 / 1e9)
for rec in sorted(sl.trace, key=lambda r: r[3])[:60]:
    print("%s slab %2d slot %d  %8.1f -> %8.1f ms (%.1f)" % (rec[0], rec[1], rec[2], 1e3*rec[3], 1e3*rec[4], 1e3*(rec[4]-rec[3])))
