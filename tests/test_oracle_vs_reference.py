"""CPU: pins the oracle (oracle/metada_oracle.c) to the REFERENCE ITSELF.

tests/golden/*.npz hold outputs of the reference's own, unmodified LETKF.hpp / ETKF.hpp / EnKF.hpp /
IdentityObsOperator.hpp / Location.hpp / Ensemble.hpp compiled in oracle/_ref (Eigen replaced by the
API shim in oracle/eigen_shim; generator: tests/golden/make_goldens.py) on the reference's tutorial
data and on a seeded synthetic case.  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

from oracle import orc
from tests.common import analysis_errors, rel_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tutorial_36x18.npz", "synthetic_23x17.npz"]


def load(name):
    g = np.load(os.path.join(G, name))
    d = {k: g[k] for k in g.files}
    # the reference narrows config reals to float (ConfigValue.hpp:108; GridObservation.hpp:244)
    d["err"] = np.full(int(d["P"]), float(np.float32(d["err_cfg"])))
    d["valid"] = np.ones(int(d["P"]), np.uint8)
    return d


@pytest.mark.parametrize("name", CASES)
def test_hx_bit_exact_vs_reference_idw4(name):
    g = load(name)
    for m in range(int(g["k"])):
        h = orc.hx_idw4(g["X"][m], g["ox"], g["oy"], g["oz"])
        assert np.array_equal(h, g["HX"][m])          # IdentityObsOperator.hpp:154-180, 594-676
    assert np.array_equal(g["err"] ** 2, g["var"])      # getCovariance = error^2, float-narrowed error


def test_tutorial_known_values_from_survey():
    g = load("tutorial_36x18.npz")
    assert g["HX"][0][0] == 0.81915203999971009
    c = g["counts"]
    assert (c.min(), c.max(), int(c.sum())) == (3, 16, 6252)
    assert abs(g["mean_letkf_snapshot"].sum() - (-222.303687100821)) < 1e-8
    assert abs(g["mean_letkf"].sum() - (-222.448687770419)) < 1e-8
    assert abs(np.abs(g["Xa_letkf"] - g["Xa_letkf_snapshot"]).max() - 1.569e-2) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_selection_counts_and_mean_bit_exact_vs_reference(name):
    g = load(name)
    counts = orc.select_counts(int(g["nx"]), int(g["ny"]), g["ox"], g["oy"], float(g["radius"]))
    assert np.array_equal(counts, g["counts"])          # Location::distance_to <= r, LETKF.hpp:159-165
    assert np.array_equal(orc.ensemble_mean(g["X"])[0], g["mean_b"])   # Ensemble.hpp:105-114


@pytest.mark.parametrize("name", CASES)
def test_letkf_as_written_matches_reference(name):
    g = load(name)
    r = orc.letkf(g["X"], g["ox"], g["oy"], g["oz"], g["yo"], g["err"], radius=float(g["radius"]),
                  inflation=float(g["inflation"]), mode=orc.MODE_REF_COMPAT, semantics=orc.SEM_AS_WRITTEN)
    assert rel_err(r["Xa"], g["Xa_letkf"]) < 1e-12
    assert rel_err(r["Xa"].sum(0)[0] * (1.0 / int(g["k"])), g["mean_letkf"]) < 1e-13


@pytest.mark.parametrize("name", CASES)
def test_letkf_snapshot_matches_reference_with_cached_H(name):
    g = load(name)
    r = orc.letkf(g["X"], g["ox"], g["oy"], g["oz"], g["yo"], g["err"], radius=float(g["radius"]),
                  inflation=float(g["inflation"]), mode=orc.MODE_REF_COMPAT, semantics=orc.SEM_SNAPSHOT)
    assert rel_err(r["Xa"], g["Xa_letkf_snapshot"]) < 1e-12
    assert np.array_equal(r["counts"], g["counts"])


@pytest.mark.parametrize("name", CASES)
def test_etkf_matches_reference(name):
    g = load(name)
    Xa = orc.etkf(g["X"], g["ox"], g["oy"], g["oz"], g["yo"], g["err"], inflation=float(g["inflation"]))
    assert rel_err(Xa, g["Xa_etkf"]) < 1e-11


@pytest.mark.parametrize("name", CASES)
def test_enkf_matches_reference_with_its_own_draws(name):
    g = load(name)
    Xa, diag = orc.enkf(g["X"], g["ox"], g["oy"], g["oz"], g["yo"], g["err"], g["enkf_Z"],
                        inflation=float(g["inflation"]), want_gain_stats=True)
    assert rel_err(Xa, g["Xa_enkf"]) < 1e-12
    ref = dict(zip(("innovation_norm", "background_spread", "analysis_spread", "max_kalman_gain",
                    "min_kalman_gain", "condition_number"), g["enkf_diag"]))
    for key, v in ref.items():
        assert abs(diag[key] - v) <= 1e-12 * abs(v), (key, diag[key], v)


def test_metrics_match_reference_header_bit_exactly():
    """orc_metrics against the reference's own Metrics.hpp (tests/golden/metrics_7x11x6x3.npz, made by
    oracle/_ref/ref_metrics): same loops, same order, no contraction -> identical bits."""
    g = np.load(os.path.join(G, "metrics_7x11x6x3.npz"))
    m = orc.metrics(g["X"], g["truth"])
    assert np.array_equal(m["mean"], g["mean"])
    assert np.array_equal(m["spread"], g["spread"])
    got = np.array([m["rmse"], m["bias"], m["correlation"], m["crps"], m["avg_spread"]])
    assert np.array_equal(got, g["scalars"]), (got, g["scalars"])
    # sanity against closed forms
    X, t = g["X"], g["truth"]
    assert abs(m["rmse"] - np.sqrt(((X.mean(0) - t) ** 2).mean())) < 1e-14
    assert abs(m["avg_spread"] - X.std(0, ddof=1).mean()) < 1e-14


@pytest.mark.parametrize("name", CASES)
def test_lwenkf_matches_reference(name):
    """LWEnKF<SimpleBackendTag>::Analyse (the reference's own, unmodified LWEnKF.hpp) for every weighting scheme and
    localisation function; the likelihood weights underflow to 0 / 0 on the tutorial set in the reference too."""
    import json
    g = load(name)
    cases = json.loads(str(g["lwenkf_cases"]))
    for i, (weighting, locfn, radius) in enumerate(cases):
        want, wdiag = g[f"lwenkf{i}_Xa"], g[f"lwenkf{i}_diag"]
        Xa, diag = orc.lwenkf(g["X"], g["ox"], g["oy"], g["oz"], g["yo"], np.sqrt(g["var"]), g[f"lwenkf{i}_Z"],
                              inflation=float(g["inflation"]), radius=float(np.float32(radius)),   # ConfigValue.hpp:108 narrowing
                              loc_fn=orc.LW_LOCFN[locfn],
                              weighting=orc.LW_WEIGHTING[weighting])
        if np.isnan(want).any():
            assert np.isnan(Xa).any(), (name, weighting)
            continue
        cond = wdiag[5]
        tol = max(1e-10, 3e-15 * cond)                  # explicit inverse of S: forward error ~ cond(S) eps
        em, ep = analysis_errors(Xa, want)
        assert em < tol and ep < tol, (name, weighting, locfn, em, ep, cond)
        for j, nm in enumerate(("innovation_norm", "background_spread", "analysis_spread", "max_kalman_gain", "min_kalman_gain",
                                "condition_number", "max_weight", "min_weight", "weight_variance")):
            assert abs(diag[j] - wdiag[j]) <= max(1e-9, 3e-14 * cond) * max(abs(wdiag[j]), 1e-30) + (1e-30 if j == 8 else 0), (nm, diag[j], wdiag[j])
