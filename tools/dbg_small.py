import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from oracle import orc
from tests.common import make_case, analysis_errors
ctx = mb.Context(0)
nx, ny = int(sys.argv[1]), int(sys.argv[2])
for k in (80, 24):
    X, o = make_case(nx, ny, 2, k, int(0.44 * nx * ny), seed=k)
    ens = mb.Ensemble(ctx, nx, ny, 2, k); ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    try:
        st = capi.letkf_analyse(ens, obs, capi.make_params(5.0, 1.05, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
        print(k, st)
    except Exception as e:
        print(k, "ERR", e, ctx.last_stats()[:8])
    Xa = ens.download()
    cols = np.arange(0, nx * ny, 37)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=5.0, inflation=1.05, cols=cols)
    cy, cx = cols // nx, cols % nx
    d = np.abs(Xa[:, :, cy, cx] - ref["Xa"][:, :, cy, cx]).max(axis=(0, 1))
    print("   cols bad:", [(int(c), float(e)) for c, e in zip(cols, d) if e > 1e-9][:12], "of", len(cols))
