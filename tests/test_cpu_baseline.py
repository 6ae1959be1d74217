"""The CPU arm of bench.py (oracle/cpu_baseline.c: cell index, tridiagonal QL eigensolver, -O3 -march=native) against
the parity oracle: same canonical analysis to 1e-10, same local-observation counts."""
import numpy as np
import pytest

from oracle import orc
from tests.common import analysis_errors, make_case


@pytest.mark.parametrize("nx,ny,nz,k,P,radius,infl", [(17, 13, 3, 12, 160, 4.0, 1.0), (12, 10, 2, 40, 150, 5.0, 1.05),
                                                      (9, 8, 1, 80, 90, 3.5, 1.0)])
def test_cpu_baseline_matches_the_oracle(nx, ny, nz, k, P, radius, infl):
    X, o = make_case(nx, ny, nz, k, P, seed=7 + k, invalid_frac=0.05, out_of_grid=4)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=radius, inflation=infl)
    Xa, tot = orc.cpu_baseline_letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=radius, inflation=infl)
    em, ep = analysis_errors(Xa, ref["Xa"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    assert tot == int(ref["counts"].sum())


def test_cpu_baseline_column_subset_and_empty_columns():
    X, o = make_case(20, 16, 2, 10, 3, seed=3)
    o["x"][:] = 1; o["y"][:] = 1
    cols = np.array([0, 5, 20 * 15 + 19], np.int64)
    Xa, tot = orc.cpu_baseline_letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=2.0, inflation=1.21, cols=cols)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"], radius=2.0, inflation=1.21, cols=cols)
    assert np.abs(Xa - ref["Xa"]).max() < 1e-12
    far = Xa[:, :, 15, 19]
    m = X[:, :, 15, 19].mean(0)
    assert np.allclose(far, m + (X[:, :, 15, 19] - m) * 1.1, rtol=0, atol=1e-14)
