"""GPU vs the committed golden vectors produced by the reference's own headers (tests/golden)."""
import numpy as np
import pytest

import metada_b200 as mb
from metada_b200 import capi
from tests.common import analysis_errors, rel_err
from tests.test_oracle_vs_reference import CASES, load

pytestmark = pytest.mark.gpu


def _setup(ctx, g):
    ens = mb.Ensemble(ctx, int(g["nx"]), int(g["ny"]), 1, int(g["k"]))
    ens.upload(g["X"])
    obs = mb.Observations(ctx, g["ox"], g["oy"], g["oz"], g["yo"], g["err"], g["valid"])
    return ens, obs


@pytest.mark.parametrize("name", CASES)
def test_hx_counts_mean_bit_exact_vs_reference(ctx, name):
    g = load(name)
    ens, obs = _setup(ctx, g)
    obs.hx(ens)
    assert np.array_equal(obs.hx_download(("Y",))["Y"], g["HX"].T)
    assert np.array_equal(obs.query_counts(ens, float(g["radius"])), g["counts"])
    assert np.array_equal(ens.mean()[0], g["mean_b"])
    ens.close(); obs.close()


@pytest.mark.parametrize("name", CASES)
def test_letkf_ref_compat_vs_reference_snapshot(ctx, name):
    g = load(name)
    ens, obs = _setup(ctx, g)
    capi.letkf_analyse(ens, obs, capi.make_params(float(g["radius"]), float(g["inflation"]), mb.MODE_REF_COMPAT, 0))
    em, ep = analysis_errors(ens.download(), g["Xa_letkf_snapshot"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    assert rel_err(ens.mean()[0], g["mean_letkf_snapshot"]) < 1e-10
    ens.close(); obs.close()


@pytest.mark.parametrize("name", CASES)
def test_etkf_and_enkf_vs_reference(ctx, name):
    g = load(name)
    ens, obs = _setup(ctx, g)
    capi.etkf_analyse(ens, obs, float(g["inflation"]))
    em, ep = analysis_errors(ens.download(), g["Xa_etkf"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    ens.close(); obs.close()
    ens, obs = _setup(ctx, g)
    diag = capi.enkf_analyse(ens, obs, float(g["inflation"]), Z=g["enkf_Z"], want_gain_stats=True)
    em, ep = analysis_errors(ens.download(), g["Xa_enkf"])
    assert em < 1e-10 and ep < 1e-9, (em, ep)
    ref = dict(zip(("innovation_norm", "background_spread", "analysis_spread", "max_kalman_gain",
                    "min_kalman_gain", "condition_number"), g["enkf_diag"]))
    for key, v in ref.items():
        assert abs(diag[key] - v) <= 1e-8 * abs(v), (key, diag[key], v)
    ens.close(); obs.close()
