"""The schedule of the composite minimax polynomial iteration (tools/gen_ns_schedule.py ->
metada_b200/csrc/ns_schedule_table.h): every table entry is checked by direct evaluation, and the kernel's
iteration is emulated in NumPy (tests/ns_emul.py) against the eigen-decomposition."""
import numpy as np
import pytest

from tests import ns_emul

LD = np.longdouble


def _image(c, rho):
    u = np.cos(np.pi * np.arange(6001) / 6000).astype(LD)
    e = -LD(rho) * u
    t = np.zeros_like(e)
    for ci in c[::-1]:
        t = t * e + LD(ci)
    return (LD(1) - e) * t * t


def test_every_step_entry_contracts_as_claimed():
    steps, _ = ns_emul.load_tables()
    rho = np.array([s[0] for s in steps])
    assert np.all(np.diff(rho) < 0) and rho[0] > 0.999997 and rho[-1] < 3e-7
    for rg, rout, c, kind, togo in steps:
        p = _image(c, rg)
        got = float(max(1 - p.min(), p.max() - 1))
        # (coefficients are rounded to double; close to rho = 1 the claim is about the gap 1 - rho, whose margin is 10 %)
        assert got <= max(rout * (1 + 1e-6) + 1e-15, rout + 0.01 * (1 - rout)), (rg, kind, got, rout)
        if kind > 10:
            assert rout <= 4e-14 and togo == kind - 10
        else:
            assert ns_emul.with_margin(rout) < rg             # a stage always contracts
    # products_to_go is what following the policy costs
    for i, (rg, rout, c, kind, togo) in enumerate(steps):
        n, j = 0, i
        while True:
            _, ro, _, kd, _ = steps[j]
            if kd > 10:
                n += kd - 10
                break
            n += kd + 2
            j = ns_emul.step_index(steps, ns_emul.with_margin(ro))
            assert j > i or j >= 0
        assert n == togo, (i, n, togo)


def test_every_start_entry():
    steps, starts = ns_emul.load_tables()
    kap = np.array([s[0] for s in starts])
    assert np.all(np.diff(kap) > 0) and kap[-1] >= ns_emul.KAPPA_MAX
    for kg, rho0, a, deg, total, seq in starts:
        xi = (LD(1) + LD(kg - 1) * (np.cos(np.pi * np.arange(6001) / 6000).astype(LD) + 1) / 2)
        q = LD(a[0]) + LD(a[1]) * xi + LD(a[2]) * xi * xi
        p = xi * q * q
        assert float(max(1 - p.min(), p.max() - 1)) <= rho0 * (1 + 1e-6) + 1e-15
        j = ns_emul.step_index(steps, ns_emul.with_margin(rho0))
        assert j >= 0 and total == 1 + deg + steps[j][4]
        # the listed sequence is what following the a-priori bounds step by step gives, and it ends with a finish whose
        # design interval covers the bound carried into it
        want = [j]
        while steps[want[-1]][3] <= 10:
            want.append(ns_emul.step_index(steps, ns_emul.with_margin(steps[want[-1]][1])))
        assert seq == want and len(seq) <= 8, (kg, seq, want)
        rho = ns_emul.with_margin(rho0)
        for jj in seq:
            assert steps[jj][0] >= rho                      # the step was designed for an interval that contains rho
            rho = ns_emul.with_margin(steps[jj][1])


def _c5_like(rng, k, p, sigma):
    Y = 0.5 * rng.standard_normal((p, k))
    Y -= Y.mean(1, keepdims=True)
    w = rng.uniform(0, 1, p) ** 2 / sigma ** 2
    Yw = Y * np.sqrt(w)[:, None]
    return (k - 1.0) * np.eye(k) + Yw.T @ Yw


@pytest.mark.parametrize("k,p,sigma,max_products,tol", [
    (80, 87, 0.3, 11, 3e-14), (80, 87, 0.1, 14, 3e-14), (80, 140, 0.1, 15, 3e-14), (40, 93, 0.1, 15, 3e-14),
    (128, 90, 0.1, 14, 3e-14), (80, 87, 0.05, 16, 5e-14), (80, 87, 0.03, 18, 1e-13), (80, 87, 0.02, 20, 2e-13),
    (24, 5, 0.1, 14, 3e-14), (80, 30, 1.0, 10, 3e-14),
    # accurate observations (round 2: the packed kernel's default limit went from a condition bound of 2000 to 1e5,
    # the table to 3e5: up there the symmetric-tile iteration agrees with numpy's eigh to ~cond * eps, and its
    # residual Z A Z - I in long double is as small as that of the eigen-decomposition itself -- condition bounds
    # 7e3, 3e4, 7e4, 7e4, 5e4; the emulation runs with kappa_max = 1e5)
    (80, 87, 0.01, 21, 1e-12), (80, 87, 0.005, 23, 2e-12), (80, 87, 0.003, 25, 5e-12),
    (128, 90, 0.003, 25, 1e-11), (40, 93, 0.004, 25, 1.5e-11)])
def test_emulated_kernel_iteration_matches_the_eigendecomposition(k, p, sigma, max_products, tol):
    rng = np.random.default_rng(k * 1000 + p)
    tables = ns_emul.load_tables()
    for _ in range(4):
        A = _c5_like(rng, k, p, sigma)
        Z, nprod, trace = ns_emul.inverse_sqrt(A, k - 1.0, tables, kappa_max=1e5)
        assert Z is not None, trace
        w, V = np.linalg.eigh(A)
        ref = (V / np.sqrt(w)) @ V.T
        assert np.abs(Z - ref).max() / np.abs(ref).max() < tol, trace
        assert nprod <= max_products, (nprod, trace)


def test_condition_bound_beyond_the_limit_is_refused():
    rng = np.random.default_rng(5)
    A = _c5_like(rng, 80, 87, 0.0008)                      # cond ~ 8e5 > 1e5
    Z, _, why = ns_emul.inverse_sqrt(A, 79.0)
    assert Z is None and why == "kappa"
    A = _c5_like(rng, 80, 87, 0.004)                       # a caller's own, lower limit (mdc_letkf_params.kappa_max)
    Z, _, why = ns_emul.inverse_sqrt(A, 79.0, kappa_max=2000.0)
    assert Z is None and why == "kappa"
