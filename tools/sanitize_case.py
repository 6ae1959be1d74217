"""Small canonical LETKF runs for compute-sanitizer (memcheck / racecheck): every column kernel once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import metada_b200 as mb
from metada_b200 import capi
from common import make_case

ctx = mb.Context(0)
for k, nz, rv, solver, err in ((80, 3, 0.0, 0, None), (40, 4, 2.0, 0, None), (104, 2, 0.0, 0, None), (48, 2, 0.0, 0, 0.01),
                               (32, 2, 0.0, 3, None), (24, 2, 0.0, 1, None)):
    X, o = make_case(9, 8, nz, k, 70, seed=k)
    if err:
        o["err"][:20] = err
    ens = mb.Ensemble(ctx, 9, 8, nz, k)
    ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(3.0, 1.02, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=rv, solver=solver))
    print(k, nz, rv, solver, st["columns"], st["max_sweeps"], st["redo_transforms"], st["numeric_failures"], flush=True)
    ens.close(); obs.close()
ctx.close()
