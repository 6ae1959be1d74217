"""One LETKF analysis of a small C5-like case (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
k = int(sys.argv[2]) if len(sys.argv) > 2 else 80
nz = int(sys.argv[3]) if len(sys.argv) > 3 else 60
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 2
P = int(round(0.4444 * n * n))
ctx = mb.Context(0)
ens = mb.Ensemble(ctx, n, n, nz, k)
o = syn.observations(P, n, n, nz, seed=42)
for it in range(2):
    ens.fill_synthetic(1000)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    st = capi.letkf_analyse(ens, obs, capi.make_params(8.0, 1.0, mode, 1))
    obs.close()
print(st)
