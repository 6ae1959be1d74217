"""Accuracy of the three canonical solvers on the mixed-conditioning test case (vs the oracle)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import metada_b200 as mb
from metada_b200 import capi
from common import make_case, analysis_errors
from oracle import orc

ctx = mb.Context(0)
for k, rv, err in ((48, 0.0, 0.01), (40, 1.5, 0.01), (48, 0.0, 0.003), (48, 0.0, 0.03)):
    nx, ny, nz = 14, 13, 3
    X, o = make_case(nx, ny, nz, k, 160, seed=77 + k)
    corner = (o["x"] < 6) & (o["y"] < 6)
    o["err"][corner] = err
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=3.0, radius_v=rv)
    for solver in (1, 2, 3):
        ens = mb.Ensemble(ctx, nx, ny, nz, k)
        ens.upload(X)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        st = capi.letkf_analyse(ens, obs, capi.make_params(3.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv, solver=solver))
        em, ep = analysis_errors(ens.download(), ref["Xa"])
        print(k, rv, err, "solver", solver, "em %.2e ep %.2e" % (em, ep), "redo", st["redo_transforms"], "maxit", st["max_sweeps"], flush=True)
        ens.close(); obs.close()
