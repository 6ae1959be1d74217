"""CPU checks of the oracle's GEOGRAPHIC / multi-variable restatement.

Pins: orc_distance_geo against the reference's own Location.hpp (tests/golden/location_geographic.npz, made by
oracle/_ref/ref_location from the unmodified header) -- bit-exact, NaNs at antipodes included; orc_geo_locate + the
per-variable H against the reference's own IdentityObsOperator.hpp driven through mock WRF-type backends
(tests/golden/obsop_geographic.npz, made by oracle/_ref/ref_obsop_geo) -- bit-exact.  An independent NumPy
restatement cross-checks both."""
import os

import numpy as np
import pytest

from metada_b200 import synthetic as syn
from oracle import orc
from tests import np_twin
from tests.common import rel_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_haversine_matches_reference_location_header_bit_exactly():
    g = np.load(os.path.join(G, "location_geographic.npz"))
    got = np.array([orc.distance_geo(*r) for r in g["pairs"]])
    same = (got == g["km"]) | (np.isnan(got) & np.isnan(g["km"]))
    assert same.all(), np.nonzero(~same)[0][:10]
    assert np.isnan(g["km"]).sum() > 0          # the reference returns NaN for (near-)antipodal pairs: a > 1
    # closed-form sanity: one degree of latitude on a 6371 km sphere
    assert abs(orc.distance_geo(10.0, 20.0, 11.0, 20.0) - 6371.0 * np.pi / 180.0) < 1e-9
    assert orc.distance_geo(10.0, 179.5, 10.0, -179.5) < 120.0   # across the dateline


def test_cartesian_distance_matches_reference_location_header_bit_exactly():
    """Location::distance_to for CARTESIAN locations (Location.hpp:217-225), the third coordinate system: restated and
    pinned to the reference's own header.  (No filter reaches it in the reference -- H throws for a CARTESIAN
    observation, IdentityObsOperator.hpp:251-255 -> Location.hpp:100-103 -- hence no device path)"""
    g = np.load(os.path.join(G, "location_geographic.npz"))
    got = np.array([orc.distance_cartesian(*r) for r in g["cart_pairs"]])
    assert np.array_equal(got, g["cart_dist"])
    assert (g["cart_dist"][800:850] == 0.0).all() and (g["cart_dist"][850:900] > 0.0).all()
    assert orc.distance_cartesian(1.0, 2.0, 3.0, 4.0, 6.0, 15.0) == 13.0


@pytest.mark.parametrize("fname", ["obsop_geographic.npz", "obsop_geographic_wstag.npz"])
def test_locate_and_variable_h_match_reference_obs_operator_bit_exactly(fname):
    """IdentityObsOperator::apply on GEOGRAPHIC observations of a [5, 5, 1]-level three-variable state: nearest grid
    point and level (:484-530), 4-of-8 IDW (:594-638) in the observation's own variable (:681-711), invalid -> 0.
    wstag: the second variable is staggered in the vertical (6 levels on a 5-level geometry, WRF's W)."""
    g = np.load(os.path.join(G, fname))
    ox, oy, oz = orc.geo_locate(g["olat"], g["olon"], g["olev"], g["lat"], g["lon"], g["vc"])
    h = orc.hx_ext(g["state"], ox, oy, oz, g["var_nlev"], g["ovar"], g["valid"])
    assert np.array_equal(h, g["HX"])
    assert (g["HX"][g["valid"] == 0] == 0.0).all() and len(np.unique(oz)) == 5
    # the first five observations sit exactly on grid points (3, 4..8): H is that point's value to 1e-11
    # (weight 1e12 against three neighbours of weight <= 1)
    for i in range(5):
        if g["valid"][i]:
            v, off = int(g["ovar"][i]), int(np.concatenate([[0], np.cumsum(g["var_nlev"])])[g["ovar"][i]])
            lev = off + (int(oz[i]) if g["var_nlev"][v] > 1 else 0)
            assert (ox[i], oy[i]) == (4 + i, 3)
            assert abs(h[i] - g["state"][lev, 3, 4 + i]) < 1e-10


def _np_locate(olat, olon, olev, glat, glon, vc):
    d = np.sqrt((olon[:, None] - glon.ravel()[None, :]) ** 2 + (olat[:, None] - glat.ravel()[None, :]) ** 2)
    idx = d.argmin(1)                            # numpy argmin = first minimum, as the strict '<' scan
    nx = glat.shape[1]
    z = np.abs(olev[:, None] - vc[None, :]).argmin(1) if vc is not None else np.zeros(len(olat), int)
    return idx % nx, idx // nx, z


def test_geo_locate_first_minimum_and_levels():
    lat, lon = syn.geography(31, 23)
    vc = np.array([1000.0, 925.0, 850.0, 700.0, 500.0, 300.0])
    o = syn.geo_observations(400, lat, lon, vc, seed=5)
    ox, oy, oz = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    ex, ey, ez = _np_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    assert np.array_equal(ox, ex) and np.array_equal(oy, ey) and np.array_equal(oz, ez)
    # exact ties (regular grid with representable spacing, observations on cell corners / edge midpoints):
    # the FIRST grid point in linear order wins
    lat, lon = syn.geography(9, 7, lat0=10.0, lon0=20.0, dlat=0.25, dlon=0.5, curvilinear=False)
    olat = np.array([10.125, 10.125, 10.25, 11.5 + 0.125])
    olon = np.array([20.25, 20.5, 20.75, 24.0 + 0.25])
    ox, oy, oz = orc.geo_locate(olat, olon, None, lat, lon, None)
    assert list(zip(ox, oy)) == [(0, 0), (1, 0), (1, 1), (8, 6)]
    assert (oz == 0).all()


def _np_letkf_geo(X, o, ox, oy, lat, lon, radius, var_nlev=None, ovar=None, oz=None, Xobs=None):
    """Independent NumPy restatement: canonical transform per column (numpy.linalg.eigh), haversine selection,
    Gaspari-Cohn weights, H by 4-point IDW at the located integer coordinates (exact hit: weight 1e12)."""
    Xa_grid = X                                     # the ensemble that is analysed (its own column set lat / lon)
    if Xobs is not None:
        X = Xobs                                    # staggered grids: H reads the ensemble that holds the variables
    k, nz, ny, nx = X.shape
    var_nlev = [nz] if var_nlev is None else list(var_nlev)
    off = np.concatenate([[0], np.cumsum(var_nlev)])
    nzg = max(var_nlev)
    P = len(ox)
    Y = np.empty((P, k))
    for i in range(P):
        v = 0 if ovar is None else int(ovar[i])
        x = float(np.clip(ox[i], 0, nx - 1)); y = float(np.clip(oy[i], 0, ny - 1))
        z = float(np.clip(oz[i] if oz is not None else 0, 0, nzg - 1))
        i0, j0, k0 = int(x), int(y), int(z)
        i1, j1, k1 = min(i0 + 1, nx - 1), min(j0 + 1, ny - 1), min(k0 + 1, nzg - 1)
        if nzg == 1:
            nb = [(i0, j0, 0), (i1, j0, 0), (i0, j1, 0), (i1, j1, 0)]
        else:
            c8 = [(i0, j0, k0), (i1, j0, k0), (i0, j1, k0), (i1, j1, k0), (i0, j0, k1), (i1, j0, k1), (i0, j1, k1), (i1, j1, k1)]
            nb = sorted(c8, key=lambda c: np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2))[:4]   # stable
        ws = wsum = 0.0
        for (ii, jj, kk) in nb:
            dd = np.sqrt((x - ii) ** 2 + (y - jj) ** 2 + (z - kk) ** 2)
            w = 1e12 if dd == 0.0 else 1.0 / dd
            lev = off[v] + (kk if var_nlev[v] > 1 else 0)
            ws = ws + w * X[:, lev, jj, ii]
            wsum += w
        Y[i] = ws / wsum
    Yp = Y - Y.mean(1, keepdims=True)
    d = o["value"] - Y.mean(1)
    X = Xa_grid
    k, nz, ny, nx = X.shape
    Xa = X.copy()
    for gy in range(ny):
        for gx in range(nx):
            dist = np.array([orc.distance_geo(lat[gy, gx], lon[gy, gx], a, b) for a, b in zip(o["lat"], o["lon"])])
            idx = np.nonzero(dist <= radius)[0]
            if len(idx) == 0:
                continue
            rho = np.array([np_twin.gaspari_cohn(dd / (0.5 * radius)) for dd in dist[idx]])
            rinv = rho / o["err"][idx] ** 2
            Yl = Yp[idx]
            A = (Yl.T * rinv) @ Yl + (k - 1) * np.eye(k)
            ev, V = np.linalg.eigh(A)
            wa = V @ ((V.T @ ((Yl.T * rinv) @ d[idx])) / ev)
            Wa = (V * np.sqrt((k - 1) / ev)) @ V.T
            x = X[:, :, gy, gx]                       # [k, nz]
            m = x.mean(0)
            Xa[:, :, gy, gx] = m[None, :] + (wa[:, None] + Wa).T @ (x - m[None, :])   # Xa = xbar + X'(w 1^T + W)
    return Xa


def test_letkf_ext_geographic_multivariable_against_numpy():
    nx, ny, k = 12, 9, 10
    var_nlev = [3, 3, 1]
    nz = sum(var_nlev)
    lat, lon = syn.geography(nx, ny, lat0=44.0, lon0=170.0, dlat=0.4, dlon=1.1)   # crosses the dateline
    assert lon.min() < -170 and lon.max() > 170
    vc = np.array([1000.0, 850.0, 500.0])
    o = syn.geo_observations(60, lat, lon, vc, seed=11)
    rng = np.random.default_rng(1)
    ovar = rng.integers(0, 3, 60).astype(np.int32)
    X = syn.ensemble(k, nx, ny, nz, seed=1234)
    ox, oy, oz = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    radius = 150.0
    r = orc.letkf_ext(X, ox, oy, oz, o["value"], o["err"], o["valid"], radius=radius, glat=lat, glon=lon,
                      olat=o["lat"], olon=o["lon"], var_nlev=var_nlev, ovar=ovar)
    counts, margin = orc.select_counts_geo(lat, lon, o["lat"], o["lon"], radius)
    assert np.array_equal(r["counts"], counts) and counts.max() > 3 and margin > 1e-9
    ref = _np_letkf_geo(X, o, ox, oy, lat, lon, radius, var_nlev, ovar, oz)
    assert rel_err(r["Xa"], ref) < 1e-11
    assert np.abs(r["Xa"] - X).max() > 1e-3      # the analysis did something


def test_letkf_ext_without_extensions_is_orc_letkf():
    X = syn.ensemble(8, 10, 7, 2, seed=77)
    o = syn.observations(40, 10, 7, 2, seed=5)
    a = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=3.0, radius_v=1.0)
    b = orc.letkf_ext(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=3.0, radius_v=1.0)
    assert np.array_equal(a["Xa"], b["Xa"]) and np.array_equal(a["counts"], b["counts"])
    # one declared variable spanning every level is the same thing
    c = orc.letkf_ext(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=3.0, radius_v=1.0,
                      var_nlev=[2], ovar=np.zeros(40, np.int32))
    assert np.array_equal(a["Xa"], c["Xa"])


def _staggered_case(k=10):
    """Mass grid 16 x 12 with variables T, QV ([3, 3] levels); U grid 17 x 12 (one more column, half a cell to the
    west) holding the single variable U; observations of T and QV."""
    nx, ny, P = 16, 12, 120
    lat, lon = syn.geography(nx, ny, lat0=40.0, lon0=5.0, dlat=0.3, dlon=0.4)
    j, i = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx + 1, dtype=np.float64) - 0.5, indexing="ij")
    ulat = 40.0 + 0.3 * j + 0.004 * np.sin(i / 7.0)
    ulon = 5.0 + 0.4 * i * (1.0 + 0.002 * j) + 0.003 * np.cos(j / 5.0)
    vc = np.array([1000.0, 850.0, 500.0])
    o = syn.geo_observations(P, lat, lon, vc, seed=21)
    ovar = np.random.default_rng(2).integers(0, 2, P).astype(np.int32)
    Xm = syn.ensemble(k, nx, ny, 6, seed=400)
    Xu = syn.ensemble(k, nx + 1, ny, 3, seed=401)
    return lat, lon, ulat, ulon, vc, o, ovar, Xm, Xu


def test_staggered_grid_analysis_reads_h_from_the_mass_grid():
    lat, lon, ulat, ulon, vc, o, ovar, Xm, Xu = _staggered_case()
    ox, oy, oz = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    kw = dict(radius=120.0, olat=o["lat"], olon=o["lon"], var_nlev=[3, 3], ovar=ovar)
    r = orc.letkf_ext(Xu, ox, oy, oz, o["value"], o["err"], o["valid"], glat=ulat, glon=ulon, Xobs=Xm, **kw)
    ref = _np_letkf_geo(Xu, o, ox, oy, ulat, ulon, 120.0, [3, 3], ovar, oz, Xobs=Xm)
    assert rel_err(r["Xa"], ref) < 1e-11 and np.abs(r["Xa"] - Xu).max() > 1e-3
    counts, _ = orc.select_counts_geo(ulat, ulon, o["lat"], o["lon"], 120.0)
    assert np.array_equal(r["counts"], counts)
    # H on the analysed ensemble itself through the same entry is the ordinary analysis
    a = orc.letkf_ext(Xm, ox, oy, oz, o["value"], o["err"], o["valid"], glat=lat, glon=lon, **kw)
    b = orc.letkf_ext(Xm, ox, oy, oz, o["value"], o["err"], o["valid"], glat=lat, glon=lon, Xobs=Xm, **kw)
    assert np.array_equal(a["Xa"], b["Xa"])
