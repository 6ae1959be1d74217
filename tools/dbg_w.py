"""Per-column transform W of each canonical solver against the oracle (mixed-conditioning case)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import metada_b200 as mb
from metada_b200 import capi
from common import make_case
from oracle import orc

ctx = mb.Context(0)
k, nx, ny, nz = 48, 14, 13, 3
X, o = make_case(nx, ny, nz, k, 160, seed=77 + k)
corner = (o["x"] < 6) & (o["y"] < 6)
o["err"][corner] = 0.01
ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=3.0, want_W=True)
for (gx, gy) in ((0, 0), (2, 2), (5, 5), (10, 10)):
    col = gy * nx + gx
    Wr = ref["W"][col]
    for solver in (1, 2, 3):
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        ens.upload(X)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        W = capi.letkf_column_transform(ens, obs, capi.make_params(3.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, solver=solver), col)
        wm_r, wm = Wr.mean(1), W.mean(1)      # ~ w (rows of Z sum to a constant)
        Zr, Z = Wr - Wr.mean(1, keepdims=True), W - W.mean(1, keepdims=True)
        print((gx, gy), "solver", solver, "W %.1e" % (np.abs(W - Wr).max() / np.abs(Wr).max()),
              "rowmean %.1e" % (np.abs(wm - wm_r).max() / np.abs(wm_r).max()),
              "centred %.1e" % (np.abs(Z - Zr).max() / np.abs(Zr).max()),
              "asym %.1e" % (np.abs(Z - Z.T).max() / np.abs(Z).max()), flush=True)
        ens.close(); obs.close()
