"""GPU parity for GEOGRAPHIC observations and multi-variable states, through the C ABI.

Bars: nearest grid point / level of every observation bit-exact (IdentityObsOperator.hpp:484-530); local-observation
counts and sets bit-exact against the brute-force haversine scan of the oracle (LETKF.hpp:159-165 with
Location.hpp:213-217) -- the device's sin / cos / atan2 are not glibc's, so each test also shows that no pair sits
within 1e-9 km of the cutoff; Y, Y', d bit-exact; analysis mean and perturbations relative <= 1e-10.
"""
import numpy as np
import pytest

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from oracle import orc
from tests.common import analysis_errors

pytestmark = pytest.mark.gpu
TOL = 1e-10
VC = np.array([1000.0, 925.0, 850.0, 700.0, 500.0, 300.0])


def _geo_case(nx, ny, nz, k, P, seed, vc=None, **geo):
    lat, lon = syn.geography(nx, ny, **geo)
    o = syn.geo_observations(P, lat, lon, vc, seed=seed)
    X = syn.ensemble(k, nx, ny, nz, seed=900 + seed)
    return lat, lon, o, X


def _setup(ctx, X, lat, lon, o, vc=None):
    k, nz, ny, nx = X.shape
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(X)
    ens.set_geography(lat, lon, vc)
    obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
    return ens, obs


@pytest.mark.parametrize("brute", [False, True])
def test_locate_bit_exact_with_ties_and_levels(ctx, brute, monkeypatch):
    """brute = False: ring walk over the bucketed grid points (the default); True: the reference's O(P G) scan."""
    if brute:
        monkeypatch.setenv("MDC_GEO_LOCATE_BRUTE", "1")
    else:
        monkeypatch.delenv("MDC_GEO_LOCATE_BRUTE", raising=False)
    lat, lon = syn.geography(37, 29)
    o = syn.geo_observations(3000, lat, lon, VC, seed=1, margin=0.35)     # a third of them outside the domain
    X = syn.ensemble(4, 37, 29, 6, seed=901)
    ens, obs = _setup(ctx, X, lat, lon, o, VC)
    obs.locate(ens)
    x, y, z = obs.grid_coords()
    ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, VC)
    assert np.array_equal(x, ex) and np.array_equal(y, ey) and np.array_equal(z, ez)
    ens.close(); obs.close()
    # exact ties on a regular grid with representable spacing: the first grid point in linear order wins;
    # more grid points (2 x 2048 + ...) than one shared-memory tile
    lat, lon = syn.geography(96, 50, lat0=10.0, lon0=20.0, dlat=0.25, dlon=0.5, curvilinear=False)
    rng = np.random.default_rng(3)
    olat = 10.0 + 0.125 * rng.integers(-2, 2 * 50 + 2, 2000)
    olon = 20.0 + 0.25 * rng.integers(-2, 2 * 96 + 2, 2000)
    ens = mb.Ensemble(ctx, 96, 50, 1, 2)
    ens.set_geography(lat, lon)
    obs = mb.Observations.geographic(ctx, olat, olon, None, np.zeros(2000), np.ones(2000))
    obs.locate(ens)
    x, y, z = obs.grid_coords()
    ex, ey, ez = orc.geo_locate(olat, olon, None, lat, lon, None)
    assert np.array_equal(x, ex) and np.array_equal(y, ey) and (z == 0).all()
    ens.close(); obs.close()
    # a domain across the dateline (raw longitudes jump from +180 to -180) and a single-row grid
    for nx, ny, kw in ((45, 31, {"lon0": 176.5}), (40, 1, {"curvilinear": False})):
        lat, lon = syn.geography(nx, ny, **kw)
        o = syn.geo_observations(1500, lat, lon, None, seed=2, margin=0.2)
        ens = mb.Ensemble(ctx, nx, ny, 1, 2)
        ens.set_geography(lat, lon)
        obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], None, o["value"], o["err"])
        obs.locate(ens)
        x, y, z = obs.grid_coords()
        ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], None, lat, lon, None)
        assert np.array_equal(x, ex) and np.array_equal(y, ey)
        ens.close(); obs.close()


@pytest.mark.parametrize("fname", ["obsop_geographic.npz", "obsop_geographic_wstag.npz"])
def test_device_h_matches_reference_obs_operator_golden_bit_exactly(ctx, fname):
    """tests/golden/obsop_geographic*.npz: output of the reference's own IdentityObsOperator.hpp (mock WRF-type
    backends, oracle/_ref/ref_obsop_geo) -- the device's location + per-variable H reproduce it bit for bit, also
    with a variable staggered in the vertical (wstag: 6 levels on a 5-level geometry)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname))
    nz, ny, nx = g["state"].shape
    ens = mb.Ensemble(ctx, nx, ny, nz, 1)
    ens.upload(g["state"][None])
    ens.set_geography(g["lat"], g["lon"], g["vc"])
    ens.set_variables(g["var_nlev"])
    P = len(g["olat"])
    obs = mb.Observations.geographic(ctx, g["olat"], g["olon"], g["olev"], np.zeros(P), np.ones(P), g["valid"])
    obs.set_variables(g["ovar"])
    obs.hx(ens)
    assert np.array_equal(obs.hx_download(("Y",))["Y"][:, 0], g["HX"])
    ens.close(); obs.close()


@pytest.mark.parametrize("radius", [0.0, 12.0, 45.0, 110.0])
@pytest.mark.parametrize("lon0", [-104.0, 176.5])
def test_haversine_selection_counts_and_sets_bit_exact(ctx, radius, lon0):
    """lon0 = 176.5: the domain crosses the dateline (columns at +179.9 and -179.9 degrees are neighbours)."""
    lat, lon, o, X = _geo_case(45, 31, 1, 2, 1500, seed=2, lon0=lon0)
    if lon0 > 0:
        assert lon.min() < -175 and lon.max() > 175
    ens, obs = _setup(ctx, X, lat, lon, o)
    counts = obs.query_counts(ens, radius)
    ref, margin = orc.select_counts_geo(lat, lon, o["lat"], o["lon"], radius)
    assert margin > 1e-9, margin                  # no pair within rounding of the cutoff: the sets are well defined
    assert np.array_equal(counts, ref)
    if radius > 40:
        assert ref.max() > 20
    cols = np.array([0, 44, 45 * 15 + 20, 45 * 31 - 1, 777], np.int64)
    lists, cnt = obs.query_lists(ens, radius, cols, cap=1500)
    for c, lst, n in zip(cols, lists, cnt):
        want = orc.select_local_geo(lat.ravel()[c], lon.ravel()[c], o["lat"], o["lon"], radius)
        assert n == len(want)
        assert sorted(lst.tolist()) == want.tolist()
    ens.close(); obs.close()


def _check_analysis(ctx, X, lat, lon, o, vc, radius, var_nlev=None, ovar=None, solver=mb.SOLVER_AUTO, radius_v=0.0,
                    loc=mb.LOC_GASPARI_COHN, inflation=1.0, mode=mb.MODE_CANONICAL):
    ens, obs = _setup(ctx, X, lat, lon, o, vc)
    if var_nlev is not None:
        ens.set_variables(var_nlev)
        obs.set_variables(ovar)
    obs.hx(ens)
    x, y, z = obs.grid_coords()
    ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    assert np.array_equal(x, ex) and np.array_equal(y, ey) and np.array_equal(z, ez)
    st = capi.letkf_analyse(ens, obs, capi.make_params(radius, inflation, mode, loc, solver=solver, radius_v=radius_v))
    Xa = ens.download()
    ref = orc.letkf_ext(X, ex, ey, ez, o["value"], o["err"], o["valid"], radius=radius, glat=lat, glon=lon,
                        olat=o["lat"], olon=o["lon"], var_nlev=var_nlev, ovar=ovar, radius_v=radius_v, loc=loc,
                        inflation=inflation, mode=mode)
    if radius_v == 0.0:      # (with per-level transforms the device counts the level-0 sets)
        assert st["sum_local_obs"] == int(ref["counts"].sum())
        assert st["max_local_obs"] == int(ref["counts"].max())
    em, ep = analysis_errors(Xa, ref["Xa"])
    assert em < TOL and ep < TOL, (em, ep)
    assert np.abs(ref["Xa"] - X).max() > 1e-3
    ens.close(); obs.close()
    return st


@pytest.mark.parametrize("k,solver", [(12, mb.SOLVER_AUTO), (40, mb.SOLVER_AUTO), (40, mb.SOLVER_JACOBI), (80, mb.SOLVER_AUTO)])
def test_geographic_letkf_matches_oracle(ctx, k, solver):
    """k = 12: Jacobi kernel; k = 40, 80: packed Newton-Schulz kernel (+ the observation-space kernel for the columns
    with few local observations at the domain edge)."""
    lat, lon, o, X = _geo_case(30, 22, 3, k, 900, seed=4, vc=VC[:3])
    st = _check_analysis(ctx, X, lat, lon, o, VC[:3], radius=55.0, solver=solver)
    assert st["columns"] == 30 * 22


@pytest.mark.parametrize("mode", [mb.MODE_REF_COMPAT, mb.MODE_REF_ETKF])
@pytest.mark.parametrize("k,lon0", [(9, -104.0), (40, 177.8)])
def test_geographic_letkf_in_the_reference_arithmetic(ctx, mode, k, lon0):
    """The reference's own point update (LETKF.hpp:209-238: R = I, cut-off selection, explicit inverse, Cholesky
    factor, per-member scale factors -- REF_COMPAT; and its ETKF-style variant) on GEOGRAPHIC observations: what
    `letkf` computes on a WRF-shaped case, selection by haversine kilometres (Location.hpp:349-357)."""
    lat, lon, o, X = _geo_case(26, 19, 2, k, 600, seed=40 + k, vc=VC[:2], lon0=lon0)
    st = _check_analysis(ctx, X, lat, lon, o, VC[:2], radius=48.0, loc=mb.LOC_CUTOFF, inflation=1.1, mode=mode)
    assert st["columns"] == 26 * 19 and st["numeric_failures"] == 0


def test_multivariable_state_in_the_reference_arithmetic(ctx):
    """REF_COMPAT on a three-variable state with per-observation variables (geographic locations)."""
    var_nlev = [3, 4, 1]
    lat, lon, o, X = _geo_case(20, 16, sum(var_nlev), 16, 500, seed=61, vc=VC[:3])
    ovar = np.random.default_rng(62).integers(0, 3, 500).astype(np.int32)
    _check_analysis(ctx, X, lat, lon, o, VC[:3], radius=50.0, var_nlev=var_nlev, ovar=ovar, loc=mb.LOC_CUTOFF,
                    mode=mb.MODE_REF_COMPAT)


def test_geographic_few_local_observations_take_the_observation_space_kernel(ctx):
    """Sparse observations: most columns have p_loc <= 24 and 2 p_loc <= k, which the packed kernel hands to the
    observation-space kernel (its EXT instantiation); some columns have no local observation at all."""
    lat, lon, o, X = _geo_case(28, 21, 2, 64, 120, seed=9, vc=VC[:2])
    st = _check_analysis(ctx, X, lat, lon, o, VC[:2], radius=30.0, inflation=1.02)
    assert st["small_transforms"] > 100, st


def test_geographic_letkf_across_the_dateline_with_reference_localisation_function(ctx):
    lat, lon, o, X = _geo_case(26, 20, 2, 24, 700, seed=5, vc=VC[:2], lon0=178.6, lat0=-48.0)
    _check_analysis(ctx, X, lat, lon, o, VC[:2], radius=60.0, loc=mb.LOC_GAUSSIAN, inflation=1.05)


@pytest.mark.parametrize("radius_v", [0.0, 1.5])
def test_geographic_multivariable_state(ctx, radius_v):
    """Three variables ([nzg, nzg, 1] levels: two 3-D fields and a surface field), each observation reads its own
    variable; with vertical localisation the level distance is taken inside the variables."""
    var_nlev = [4, 4, 1]
    lat, lon, o, X = _geo_case(24, 18, sum(var_nlev), 32, 800, seed=6, vc=VC[:4])
    ovar = np.random.default_rng(8).integers(0, 3, 800).astype(np.int32)
    st = _check_analysis(ctx, X, lat, lon, o, VC[:4], radius=50.0, var_nlev=var_nlev, ovar=ovar, radius_v=radius_v)
    assert st["columns"] == 24 * 18


@pytest.mark.parametrize("radius_v", [0.0, 1.5])
def test_geographic_state_with_a_vertically_staggered_variable(ctx, radius_v):
    """WRF's W / PH (WRFState.hpp:89-93, 445-449): nz + 1 levels on an nz-level geometry.  H never reads the top level
    (the neighbour search runs on the geometry's levels), the column update transforms all nz + 1."""
    var_nlev = [4, 5, 1]
    lat, lon, o, X = _geo_case(22, 17, sum(var_nlev), 24, 700, seed=16, vc=VC[:4])
    ovar = np.random.default_rng(18).integers(0, 3, 700).astype(np.int32)
    st = _check_analysis(ctx, X, lat, lon, o, VC[:4], radius=50.0, var_nlev=var_nlev, ovar=ovar, radius_v=radius_v)
    assert st["columns"] == 22 * 17
    # a variable that is neither on the geometry's levels nor staggered by one is refused
    ens = mb.Ensemble(ctx, 5, 4, 4 + 7 + 1, 3)
    with pytest.raises(mb.MdcError):
        ens.set_variables([4, 7, 1])
    ens.close()


def test_grid_observations_on_a_multivariable_state(ctx):
    """GRID coordinates (integer distances) with per-observation variables: H and Y' bit-exact, analysis <= 1e-10."""
    nx, ny, k, P = 21, 17, 24, 500
    var_nlev = [3, 1, 3]
    nz = sum(var_nlev)
    X = syn.ensemble(k, nx, ny, nz, seed=4321)
    o = syn.observations(P, nx, ny, 3, seed=12)
    ovar = np.random.default_rng(9).integers(0, 3, P).astype(np.int32)
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(X)
    ens.set_variables(var_nlev)
    obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
    obs.set_variables(ovar)
    for radius_v in (0.0, 1.0):
        ens.upload(X)
        st = capi.letkf_analyse(ens, obs, capi.make_params(4.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN, radius_v=radius_v))
        ref = orc.letkf_ext(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=4.0,
                            var_nlev=var_nlev, ovar=ovar, radius_v=radius_v)
        if radius_v == 0.0:
            assert st["sum_local_obs"] == int(ref["counts"].sum())
        em, ep = analysis_errors(ens.download(), ref["Xa"])
        assert em < TOL and ep < TOL, (radius_v, em, ep)
    ens.close(); obs.close()


@pytest.mark.parametrize("k", [10, 32])
def test_staggered_grid_is_its_own_column_set_sharing_the_mass_grid_h(ctx, k):
    """U-staggered variable (17 x 12 columns with their own coordinates) next to a mass grid (16 x 12, variables T and
    QV): H(x) is evaluated on the mass-grid ensemble, the U columns are analysed with that Y' (mdc_hx_idw4 on one
    store, mdc_letkf_analyse on the other), then the mass grid itself."""
    from tests.test_oracle_geo import _staggered_case
    lat, lon, ulat, ulon, vc, o, ovar, Xm, Xu = _staggered_case(k)
    mass = mb.Ensemble(ctx, 16, 12, 6, k)
    mass.upload(Xm); mass.set_geography(lat, lon, vc); mass.set_variables([3, 3])
    ugrid = mb.Ensemble(ctx, 17, 12, 3, k)
    ugrid.upload(Xu); ugrid.set_geography(ulat, ulon, vc)
    obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
    obs.set_variables(ovar)
    prm = capi.make_params(120.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
    obs.hx(mass)
    st_u = capi.letkf_analyse(ugrid, obs, prm)          # Y' from the mass grid, columns and coordinates of the U grid
    obs.hx(mass)                                        # (an analysis drops Y'; the mass background is still untouched)
    st_m = capi.letkf_analyse(mass, obs, prm)
    ox, oy, oz = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    kw = dict(radius=120.0, olat=o["lat"], olon=o["lon"], var_nlev=[3, 3], ovar=ovar)
    ref_u = orc.letkf_ext(Xu, ox, oy, oz, o["value"], o["err"], o["valid"], glat=ulat, glon=ulon, Xobs=Xm, **kw)
    ref_m = orc.letkf_ext(Xm, ox, oy, oz, o["value"], o["err"], o["valid"], glat=lat, glon=lon, **kw)
    assert st_u["columns"] == 17 * 12 and st_u["sum_local_obs"] == int(ref_u["counts"].sum())
    assert st_m["columns"] == 16 * 12 and st_m["sum_local_obs"] == int(ref_m["counts"].sum())
    for got, ref in ((ugrid.download(), ref_u["Xa"]), (mass.download(), ref_m["Xa"])):
        em, ep = analysis_errors(got, ref)
        assert em < TOL and ep < TOL, (em, ep)
    mass.close(); ugrid.close(); obs.close()


def test_unsupported_geographic_requests_fail_loudly(ctx):
    lat, lon, o, X = _geo_case(12, 10, 1, 8, 50, seed=7)
    ens, obs = _setup(ctx, X, lat, lon, o)
    with pytest.raises(mb.MdcError, match="solvers"):
        capi.letkf_analyse(ens, obs, capi.make_params(50.0, 1.0, mb.MODE_CANONICAL, solver=mb.SOLVER_NEWTON_SCHULZ_FULL))
    with pytest.raises(mb.MdcError, match="pole"):
        obs.query_counts(ens, 7000.0)
    with pytest.raises(mb.MdcError, match="index"):
        obs.index_build(8)
    plain = mb.Ensemble(ctx, 12, 10, 1, 8)
    with pytest.raises(mb.MdcError, match="geography"):
        obs.locate(plain)
    plain.close(); ens.close(); obs.close()


@pytest.mark.parametrize("world,mode,multivar", [(2, mb.MODE_CANONICAL, False), (3, mb.MODE_CANONICAL, True),
                                                  (4, mb.MODE_REF_COMPAT, False)])
def test_sharded_geographic_analysis_is_bit_identical_to_one_store(ctx, world, mode, multivar):
    """Domain decomposition of a geographic analysis (VERDICT r1 missing 3), the ranks played one after the other on
    this device: every rank locates all observations on the global geography, keeps those whose nearest grid point
    lies in its slab, applies H to them and packs -- by box, with latitude / longitude / level / variable in the row
    -- what the other ranks' columns can reach; each rank then analyses its rows with own + received observations on
    its window of the geography.  Same lattice, same candidate order: the assembled analysis equals the one-store
    result bit for bit."""
    from metada_b200.parallel import GeoSlabLetkf
    var_nlev = [3, 3, 1] if multivar else None
    nx, ny, nz, k, P, radius = 27, 26, (7 if multivar else 2), 24, 700, 45.0
    vc = VC[:3] if multivar else VC[:2]
    lat, lon, o, X = _geo_case(nx, ny, nz, k, P, seed=70 + world, vc=vc, lon0=(177.9 if world == 3 else -104.0))
    ovar = np.random.default_rng(5).integers(0, 3, P).astype(np.int32) if multivar else None
    ens, obs = _setup(ctx, X, lat, lon, o, vc)
    if multivar:
        ens.set_variables(var_nlev)
        obs.set_variables(ovar)
    params = capi.make_params(radius, 1.04, mode, mb.LOC_GASPARI_COHN if mode == mb.MODE_CANONICAL else mb.LOC_CUTOFF)
    st1 = capi.letkf_analyse(ens, obs, params)
    one = ens.download()
    ens.close(); obs.close()
    oo = dict(o)
    oo["var"] = ovar
    jobs = [GeoSlabLetkf(ctx, lat, lon, vc, nz, k, r, world, radius, var_nlev) for r in range(world)]
    sends = []
    for job in jobs:
        job.ens.upload(np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :]))
        job.set_observations(oo)
        sends.append(job.pack_halo())
    assert sum(len(job.own) for job in jobs) == P
    out = np.empty_like(X)
    cols = halo = 0
    for r, job in enumerate(jobs):
        st = job.analyse(params, recv={src: sends[src][r] for src in range(world) if src != r})
        cols += st["columns"]
        halo += job.halo_rows_last
        out[:, :, job.y0:job.y1, :] = job.ens.download()[:, :, :job.y1 - job.y0, :]
    for job in jobs:
        job.close()
    assert cols == nx * ny and 0 < halo < P * (world - 1)
    assert np.array_equal(out, one)
    assert np.abs(one - X).max() > 1e-3
