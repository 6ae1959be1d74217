"""Shared seeded test cases (small enough for the CPU oracle to finish in seconds)."""
import numpy as np

from metada_b200 import synthetic as syn


def make_case(nx, ny, nz, k, P, seed=1, sigma=0.1, invalid_frac=0.0, out_of_grid=0):
    """Synthetic ensemble + obs.  out_of_grid > 0 moves that many obs outside the grid (H clamps
    them, IdentityObsOperator.hpp:598-600); invalid_frac marks some obs invalid."""
    X = syn.ensemble(k, nx, ny, nz, seed=1000 + seed)
    o = syn.observations(P, nx, ny, nz, seed=42 + seed, sigma=sigma)
    rng = np.random.default_rng(seed)
    if out_of_grid:
        idx = rng.choice(P, size=out_of_grid, replace=False)
        o["x"][idx[: out_of_grid // 2]] = nx + 2
        o["y"][idx[out_of_grid // 2:]] = -3
    if invalid_frac > 0:
        bad = rng.random(P) < invalid_frac
        o["valid"][bad] = 0
    return X, o


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def analysis_errors(Xa, Xref):
    """relative error on the analysis mean and on the perturbations."""
    ma, mr = Xa.mean(0), Xref.mean(0)
    return rel_err(ma, mr), rel_err(Xa - ma, Xref - mr)
