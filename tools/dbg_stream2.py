import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from tests.common import make_case
ctx = mb.Context(0)
nx, ny, nz, k, P, radius = 19, 43, 2, 24, 460, 4.0
X, o = make_case(nx, ny, nz, k, P, seed=43, out_of_grid=6)
params = capi.make_params(radius, 1.05, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
gid = np.arange(P, dtype=np.int64)
def run(y0, y1, obs_lo, obs_hi):
    halo = 1 if y1 < ny else 0
    ens = mb.Ensemble(ctx, nx, (y1 - y0) + halo, nz, k)
    ens.set_domain(0, y0, nx, ny, nx, y1 - y0)
    ens.upload(np.ascontiguousarray(X[:, :, y0:y1 + halo, :]))
    # H on the full domain (same Y' bits): use a full ensemble for hx, then analyse the sub-ensemble
    sel = np.where((o["y"] >= obs_lo) & (o["y"] < obs_hi))[0]
    full = mb.Ensemble(ctx, nx, ny, nz, k); full.upload(X)
    obs = mb.Observations(ctx, o["x"][sel], o["y"][sel], o["z"][sel], o["value"][sel], o["err"][sel], o["valid"][sel], gid=gid[sel])
    obs.hx(full)
    st = capi.letkf_analyse(ens, obs, params)
    out = ens.download()[:, :, :y1 - y0, :]
    col = np.array([2 * nx + 1], np.int64)   # local row 2, x = 1
    lists, cnt = obs.query_lists(ens, radius, col, cap=200)
    ens.close(); obs.close(); full.close()
    return out, lists[0], st
a, la, sta = run(0, ny, -100, 1000)
b, lb, stb = run(4, 8, 0, 12)
c, lc, stc = run(4, 8, -100, 1000)
d, ld, std = run(4, 9, -100, 1000)
print("one-shot vs slab(4..8, obs 0..12):", np.abs(a[:, :, 4:8] - b).max(), "vs slab all obs:", np.abs(a[:, :, 4:8] - c).max(), "slab5 all obs", np.abs(a[:, :, 4:9] - d).max())
print("b vs c", np.abs(b - c).max())
print(sta, stb, stc, sep="\n")
