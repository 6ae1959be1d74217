import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from metada_b200.parallel import SlabLetkf
from tests.common import make_case
ctx = mb.Context(0)
nx, ny, nz, k, P, radius = 19, 43, 2, 24, 460, 4.0
X, o = make_case(nx, ny, nz, k, P, seed=43, out_of_grid=6)
params = capi.make_params(radius, 1.05, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.upload(X)
obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
col = np.array([6 * nx + 1, 6 * nx + 3], np.int64)
l1, c1 = obs.query_lists(ens, radius, col, cap=200)
print("one-shot list", c1, l1[0].tolist())
world, slab_rows = 3, 4
job = SlabLetkf(ctx, nx, ny, nz, k, 0, world, radius)
host = np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :])
jobs = [SlabLetkf(ctx, nx, ny, nz, k, r, world, radius) for r in range(world)]
hosts = [np.ascontiguousarray(X[:, :, j.y0:j.y0 + j.ny_loc, :]) for j in jobs]
sends = [j.edge_pack([h[m].ctypes.data for m in range(k)], o) for j, h in zip(jobs, hosts)]
recv = {src: sends[src][0] for src in range(world) if src != 0 and 0 in sends[src]}
sl = mb.StreamedLetkf(0, nx, ny, nz, k, radius, slab_rows=slab_rows, slots=3, row_range=(job.y0, job.y1))
st = jobs[0].streamed_analyse(sl, [hosts[0][m].ctypes.data for m in range(k)], o, params, recv)
print("bounds", sl.bounds)
s = 1
ob = sl._obs[s]
y0, y1 = sl.bounds[s]
e2 = mb.Ensemble(sl.ctxs[s % 3], nx, y1 - y0 + 1, nz, k)
e2.set_domain(0, y0, nx, ny, nx, y1 - y0)
lc = np.array([(6 - y0) * nx + 1, (6 - y0) * nx + 3], np.int64)
l2, c2 = ob.query_lists(e2, radius, lc, cap=200)
print("slab list   ", c2, l2[0].tolist())
print("store size", ob.size())
