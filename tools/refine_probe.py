"""Accuracy of the packed kernel against the oracle as the condition number grows (development): kappa_max raised to
the table's limit, observation error swept."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from oracle import orc
from tests.common import analysis_errors, make_case
ctx = mb.Context(0)
for k in (80, 40, 128, 56):
    for sigma in (0.01, 0.005, 0.003, 0.002, 0.0015):
        nx, ny, nz = 16, 14, 3
        X, o = make_case(nx, ny, nz, k, 260, seed=500 + k, sigma=sigma)
        o["err"][:] = sigma
        ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.upload(X)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        st = capi.letkf_analyse(ens, obs, capi.make_params(5.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN,
                                                           solver=mb.SOLVER_NEWTON_SCHULZ, kappa_max=3e5))
        ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=5.0)
        em, ep = analysis_errors(ens.download(), ref["Xa"])
        print(json.dumps({"k": k, "sigma": sigma, "err_mean": em, "err_pert": ep, "redo": st["redo_transforms"], "max_products": st["max_sweeps"],
                          "fail": st["numeric_failures"]}), flush=True)
        ens.close(); obs.close()
