import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi
from oracle import orc
from tests.common import make_case, analysis_errors
ctx = mb.Context(0)
def setup(X, o):
    k, nz, ny, nx = X.shape
    ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    return ens, obs
# EnKF
X, o = make_case(25, 16, 1, 10, 120, seed=11)
ens, obs = setup(X, o)
Z = np.random.default_rng(7).standard_normal((120, 10))
diag = capi.enkf_analyse(ens, obs, 1.1, Z=Z, want_gain_stats=True)
ref, rdiag = orc.enkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], Z, inflation=1.1, want_gain_stats=True)
print("enkf", analysis_errors(ens.download(), ref), "cond", rdiag["condition_number"])
# ill-conditioned NS
for solver in (mb.SOLVER_NEWTON_SCHULZ, mb.SOLVER_NEWTON_SCHULZ_FULL, mb.SOLVER_JACOBI):
    X, o = make_case(12, 12, 4, 32, 150, seed=21, sigma=0.002)
    o["err"][:] = 0.002
    ens, obs = setup(X, o)
    p = capi.make_params(4.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=2.0, solver=solver)
    st = capi.letkf_analyse(ens, obs, p)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=4.0, radius_v=2.0)
    print("illcond solver", solver, analysis_errors(ens.download(), ref["Xa"]))
for k, radius_v in [(48, 0.0), (40, 1.5), (104, 0.0)]:
    nx, ny, nz = 14, 13, 3
    X, o = make_case(nx, ny, nz, k, 160, seed=77 + k)
    corner = (o["x"] < 6) & (o["y"] < 6)
    o["err"][corner] = 0.01
    ens, obs = setup(X, o)
    p = capi.make_params(3.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=radius_v, solver=mb.SOLVER_NEWTON_SCHULZ)
    st = capi.letkf_analyse(ens, obs, p)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=3.0, radius_v=radius_v)
    print("mixed", k, radius_v, analysis_errors(ens.download(), ref["Xa"]), st["redo_transforms"])
