"""GPU parity at the BASELINE.json configuration shapes.

C1 is analysed in full (100x100, k = 20, 1e3 obs, radius 10; REF_COMPAT and CANONICAL).  C3 / C4 / C5 do not
fit the oracle (or this host's memory) in full, so a window of the global grid is placed with
``mdc_ens_set_domain``, filled on the device from the GLOBAL synthetic field (``mdc_ens_fill_synthetic``, checked
bit for bit against the host generator on sample members), analysed at the configuration's k, level count,
observation density, radii and localisation, and a seeded set of columns (interior, the global domain's edge and
corner, the window's cut edges) is compared with ``orc.letkf(cols=...)`` on the same numbers.  The window is
analysed as a domain of its own by both sides (only the window's observations are given to either), so every
column of it is comparable.

Bars: selection counts bit-exact, analysis mean and perturbations relative <= 1e-10.
"""
import numpy as np
import pytest

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from oracle import orc
from tests.common import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("mode,loc", [(mb.MODE_REF_COMPAT, mb.LOC_CUTOFF), (mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)])
def test_c1_full(ctx, mode, loc):
    """BASELINE.json configs[0]: 100x100x1, 20 members, 1e3 obs, radius 10 -- every column against the oracle
    (LETKF.hpp:152-243 arithmetic in REF_COMPAT; Gaspari-Cohn R-localisation + symmetric square root in CANONICAL)."""
    nx = ny = 100
    k, P, radius = 20, 1000, 10.0
    X = syn.ensemble(k, nx, ny, 1, seed=1000)
    o = syn.observations(P, nx, ny, 1, seed=42)
    ens = mb.Ensemble(ctx, nx, ny, 1, k)
    ens.upload(X)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    counts = obs.query_counts(ens, radius)
    st = capi.letkf_analyse(ens, obs, capi.make_params(radius, 1.0, mode, loc))
    Xa = ens.download()
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=radius, mode=mode, loc=loc)
    assert np.array_equal(counts, ref["counts"])                          # bit-exact selection counts, all 1e4 columns
    assert st["columns"] == nx * ny and st["numeric_failures"] == 0
    assert st["sum_local_obs"] == int(ref["counts"].sum()) and st["max_local_obs"] == int(ref["counts"].max())
    ma, mr = Xa.mean(0), ref["Xa"].mean(0)
    assert rel_err(ma, mr) < TOL
    assert rel_err(Xa - ma, ref["Xa"] - mr) < TOL
    ens.close(); obs.close()


def _window_case(ctx, gnx, gny, nz, k, P, W, corner):
    """A W x W window of the gnx x gny global grid at the origin corner or the far corner (the far corner has no
    halo: the window ends at the global edge), device-filled, with the global observations that fall inside it."""
    if corner == "origin":
        gx0 = gy0 = 0
        nxl = nyl = W + 1                    # one read-only halo row / column on the high (cut) sides for H
    else:
        gx0, gy0 = gnx - W, gny - W
        nxl = nyl = W
    ens = mb.Ensemble(ctx, nxl, nyl, nz, k)
    ens.set_domain(gx0, gy0, gnx, gny, W, W)
    ens.fill_synthetic(1000)
    X = ens.download()
    for m in (0, k // 2, k - 1):             # the device generator is the host generator, bit for bit
        assert np.array_equal(X[m], syn.member(m, nxl, nyl, nz, 1000, gx0=gx0, gy0=gy0, gnx=gnx, gny=gny)), m
    o = syn.observations(P, gnx, gny, nz, seed=42)
    inside = (o["x"] >= gx0) & (o["x"] < gx0 + W) & (o["y"] >= gy0) & (o["y"] < gy0 + W)
    o = {key: np.ascontiguousarray(v[inside]) for key, v in o.items()}
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    return ens, obs, X, o, (gx0, gy0, nxl, nyl)


def _sample_columns(W, nxl, n, reach, seed):
    """n seeded columns of the owned W x W block: the four corners, points on every edge, points within `reach` of
    the edges, the rest anywhere."""
    rng = np.random.default_rng(seed)
    pts = {(0, 0), (W - 1, 0), (0, W - 1), (W - 1, W - 1)}
    for _ in range(n // 8):
        a = int(rng.integers(0, W))
        pts.update({(a, 0), (0, a), (a, W - 1), (W - 1, a)})
    for _ in range(n // 4):
        a, b = int(rng.integers(0, W)), int(rng.integers(0, reach + 1))
        pts.update({(a, b), (b, a), (a, W - 1 - b), (W - 1 - b, a)})
    while len(pts) < n:
        pts.add((int(rng.integers(0, W)), int(rng.integers(0, W))))
    pts = sorted(pts)[:10 ** 9]
    return np.array([y * nxl + x for x, y in pts], np.int64)


def _compare_window(ctx, gnx, gny, nz, k, P, W, corner, ncols, radius, radius_v=0.0, seed=7):
    ens, obs, X, o, (gx0, gy0, nxl, nyl) = _window_case(ctx, gnx, gny, nz, k, P, W, corner)
    params = capi.make_params(radius, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=radius_v)
    st = capi.letkf_analyse(ens, obs, params)
    Xa = ens.download()
    cols = _sample_columns(W, nxl, ncols, int(radius), seed)
    ref = orc.letkf(X, o["x"] - gx0, o["y"] - gy0, o["z"], o["value"], o["err"], o["valid"], radius=radius,
                    radius_v=radius_v, cols=cols)
    cy, cx = cols // nxl, cols % nxl
    got, want = Xa[:, :, cy, cx], ref["Xa"][:, :, cy, cx]
    assert not np.array_equal(want, X[:, :, cy, cx])                      # the oracle did analyse these columns
    # selection: the counts of the sampled columns, bit-exact
    lists, cnt = obs.query_lists(ens, radius, cols, cap=1024)
    assert np.array_equal(np.asarray(cnt), ref["counts"].reshape(-1)[cols])
    mg, mw = got.mean(0), want.mean(0)
    em, ep = rel_err(mg, mw), rel_err(got - mg, want - mw)
    assert em < TOL and ep < TOL, (em, ep, st)
    assert st["columns"] == W * W and st["numeric_failures"] == 0, st
    ens.close(); obs.close()
    return st


@pytest.mark.parametrize("corner", ["origin", "far"])
def test_c3_window(ctx, corner):
    """configs[2]: 400x400x50, 40 members, 1e5 obs, horizontal Gaspari-Cohn radius 7; 192 x 192 window, 512 columns."""
    st = _compare_window(ctx, 400, 400, 50, 40, 100_000, 192, corner, 512, 7.0)
    assert 80 < st["sum_local_obs"] / st["columns"] < 100                 # p_loc ~ 93 (edges lower it)


@pytest.mark.parametrize("corner", ["origin", "far"])
def test_c5_window(ctx, corner):
    """configs[4] (the benchmark): 1500x1500x60, 80 members, 1e6 obs, radius 8; 192 x 192 window with all 60
    levels, 512 columns.  This is the packed Newton-Schulz kernel (letkf_nsp_kernel) at its production shape."""
    st = _compare_window(ctx, 1500, 1500, 60, 80, 1_000_000, 192, corner, 512, 8.0)
    assert 75 < st["sum_local_obs"] / st["columns"] < 95
    assert st["max_sweeps"] > 0                                           # Newton-Schulz products were counted


def test_c4_window(ctx):
    """configs[3]: 1000x1000x60, 128 members, 5e5 obs, horizontal radius 8 and vertical radius 5 levels (one
    transform per level); 96 x 96 window, 48 columns x 60 levels = 2 880 transforms against the oracle."""
    st = _compare_window(ctx, 1000, 1000, 60, 128, 500_000, 96, "origin", 48, 8.0, radius_v=5.0)
    assert st["numeric_failures"] == 0


def test_c4_window_horizontal_only(ctx):
    """C4's ensemble size with one transform per column (k = 128 packed kernel, 16 warps, one CTA per SM)."""
    _compare_window(ctx, 1000, 1000, 60, 128, 500_000, 96, "far", 64, 8.0)
