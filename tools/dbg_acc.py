import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from oracle import orc
from tests.common import make_case, analysis_errors
ctx = mb.Context(0)
for k, P in ((20, 150), (40, 150), (80, 150), (128, 100)):
    X, o = make_case(12, 10, 2, k, P, seed=k)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=5.0, inflation=1.05, want_W=True)
    for jtol in (1e-9, 1e-13):
        ens = mb.Ensemble(ctx, 12, 10, 2, k); ens.upload(X)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        st = capi.letkf_analyse(ens, obs, capi.make_params(5.0, 1.05, 2, 1, jacobi_tol=jtol))
        Xa = ens.download()
        em, ep = analysis_errors(Xa, ref["Xa"])
        err = np.abs((Xa - Xa.mean(0)) - (ref["Xa"] - ref["Xa"].mean(0))).max(axis=(0, 1))
        worst = np.unravel_index(err.argmax(), err.shape)
        print(k, jtol, "em %.2e ep %.2e" % (em, ep), "sweeps", st["sum_sweeps"] / st["columns"], "worst col", worst, "ploc", ref["counts"][worst])
        ens.close(); obs.close()
    # W of the worst column
    col = int(worst[0] * 12 + worst[1])
    ens = mb.Ensemble(ctx, 12, 10, 2, k); ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    W = capi.letkf_column_transform(ens, obs, capi.make_params(5.0, 1.05, 2, 1), col)
    Wr = ref["W"][col]
    print("   W err %.2e" % (np.abs(W - Wr).max() / np.abs(Wr).max()), "W max", np.abs(Wr).max(), "Xp max", np.abs(X - X.mean(0)).max())
    ens.close(); obs.close()
