import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import metada_b200 as mb
ctx = mb.Context(0)
nx = ny = 256; nz = 60; k = 80
ens = mb.Ensemble(ctx, nx, ny, nz, k)
ens.fill_synthetic(1000)
n = nx * ny * nz
host = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(k)]
ptrs = [t.data_ptr() for t in host]
for it in range(3):
    t0 = time.perf_counter(); ens.download_ptrs(0, ptrs); ctx.sync(); t1 = time.perf_counter()
    ens.upload_ptrs(0, ptrs); ctx.sync(); t2 = time.perf_counter()
    print("download GB/s", n * k * 8 / (t1 - t0) / 1e9, "upload GB/s", n * k * 8 / (t2 - t1) / 1e9)
# pageable
hp = np.empty((k, n))
t0 = time.perf_counter(); ens.download_ptrs(0, [hp[m].ctypes.data for m in range(k)]); ctx.sync(); t1 = time.perf_counter()
print("pageable download GB/s", n * k * 8 / (t1 - t0) / 1e9)
t0 = time.perf_counter(); m = ens.mean(); t1 = time.perf_counter(); print("mean s", t1 - t0)
