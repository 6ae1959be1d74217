"""Debug: where does the streamed/sharded result differ from the one-shot analysis?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from metada_b200.parallel import SlabLetkf
from tests.common import make_case

world, slab_rows = int(sys.argv[1]), int(sys.argv[2])
ctx = mb.Context(0)
nx, ny, nz, k, P, radius = 19, 43, 2, 24, 460, 4.0
X, o = make_case(nx, ny, nz, k, P, seed=43, out_of_grid=6)
params = capi.make_params(radius, 1.05, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
outs = []
for rep in range(2):
    ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st1 = capi.letkf_analyse(ens, obs, params)
    outs.append(ens.download()); ens.close(); obs.close()
print("one-shot repeatable:", np.array_equal(outs[0], outs[1]), st1)
one_shot = outs[0]
jobs, hosts, sends = [], [], []
for r in range(world):
    job = SlabLetkf(ctx, nx, ny, nz, k, r, world, radius)
    host = np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :])
    jobs.append(job); hosts.append(host)
    sends.append(job.edge_pack([host[m].ctypes.data for m in range(k)], o))
res=[]
for rep in range(3):
  out = np.empty_like(X)
  for r, (job, host) in enumerate(zip(jobs, hosts)):
      recv = {src: sends[src][r] for src in range(world) if src != r and r in sends[src]}
      sl = mb.StreamedLetkf(0, nx, ny, nz, k, radius, slab_rows=slab_rows, slots=int(sys.argv[3]) if len(sys.argv) > 3 else 3, row_range=(job.y0, job.y1))
      st = job.streamed_analyse(sl, [host[m].ctypes.data for m in range(k)], o, params, recv)
      print("rank", r, st)
      sl.close()
      out[:, :, job.y0:job.y1, :] = host[:, :, :job.y1 - job.y0, :]
  hosts = [np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :]) for job in jobs]
  res.append(out.copy())
print("pipeline repeatable:", [np.array_equal(res[0], r_) for r_ in res])
out = res[0]
d = np.abs(out - one_shot)
print("max abs diff", d.max(), "rel", d.max() / np.abs(one_shot).max())
bad = np.argwhere(d.max(axis=(0, 1)) > 0)
print("differing columns (y, x):", bad[:40].tolist(), len(bad))
