#!/usr/bin/env python
"""Generates metada_b200/csrc/ns_schedule_table.h: the schedule of the composite minimax polynomial
iteration for Z = A^{-1/2} used by the packed Newton-Schulz column kernel (letkf_nsp.cuh).

State: Z (a polynomial in A) and the residual E = I - Z^2 A, spectrum(E) in [-rho, rho].
  stage  of degree d: T = sum_i c_i E^i,  Z <- Z T,  E <- I - (I - E) T^2        (d + 2 symmetric products)
  finish of degree f: Z <- Z sum_i c_i E^i                                        (f products)
  start  of degree s: Z0 = sum_i a_i (A/lmin)^i / sqrt(lmin),  E0 = I - Z0 (A Z0)  (A^2 is always formed: it gives the
                      Schatten-4 bound of the spectrum; s = 0: 1 product, s = 1: 2, s = 2: 3)
Plain Newton-Schulz is the stage d = 1 with Taylor coefficients (1, 1/2).  Here every polynomial is the MINIMAX
one for the interval the spectrum is known to lie in (equal-ripple x t(x)^2 on [1 - rho, 1 + rho], rescaled so the
image is centred on 1), and the sequence of degrees is chosen by dynamic programming over rho so that the number
of k x k products is minimal: 13-14 products where Chebyshev start + Newton-Schulz + series finish needs 17-20
(condition bounds 40-80).  Everything depends on one scalar (kappa for the start, rho afterwards), so the schedule
is two small tables; the start entry (chosen by kappa) lists the steps that follow, so the kernel reads all of a
column's coefficients at once.  The a-priori bounds are rigorous as long as spectrum(A) lies in [lmin, lmin kappa];
the kernel checks ||E||_F <= sqrt(k) rho before the finishing step.
"""
import os
import sys

import numpy as np
from scipy.optimize import linprog

TOL_RHO_OUT = 4e-14          # finish allowed when the image of the finishing polynomial is within 1 +- this
MARGIN = 1.002               # safety factor on every a-priori rho (sampling of the extrema, rounding)
MARGIN_TOP = 4000.0 / 4002.0 # ... while rho * MARGIN stays below this; closer to 1 the factor would pass 1, and what
MARGIN_GAP = 0.90            # matters is the LOWER end 1 - rho of the spectrum of Z^2 A: the gap shrinks by this instead
KAPPA_TABLE_MAX = 3e5        # condition bounds the start table covers (the kernel's default limit is 2e5)
GRID_EXTRA = 43              # rho-grid entries above the original top (kappa' - 1 = 4000): up to ~9.7e5


def with_margin(rho):
    m = rho * MARGIN
    return m if m <= MARGIN_TOP else 1.0 - (1.0 - rho) * MARGIN_GAP
TAYLOR_E = [1.0, 0.5, 0.375, 0.3125]


def _lp_minimax(Phi, target):
    """min delta s.t. |Phi c - target| <= delta (rows = sample points): linear Chebyshev approximation by LP"""
    m, n = Phi.shape
    cost = np.zeros(n + 1); cost[-1] = 1.0
    A = np.block([[Phi, -np.ones((m, 1))], [-Phi, -np.ones((m, 1))]])
    b = np.concatenate([target, -target])
    r = linprog(cost, A_ub=A, b_ub=b, bounds=[(None, None)] * n + [(0, None)], method="highs")
    if r.status != 0:
        raise RuntimeError(r.message)
    return r.x[:n], r.x[-1]


def _centre(p):
    s = 2.0 / (p.min() + p.max())
    return s, max(1 - p.min() * s, p.max() * s - 1)


def design_stage(rho, deg, npts=1201):
    """minimax t (degree deg) on x in [1-rho, 1+rho] in the sense  max |sqrt(x) t(x) - 1| -> min  (which also minimises
    the spread of x t(x)^2): coefficients in e = 1 - x (ascending), rescaled so the image of x t^2 is centred on 1,
    and the image half-width."""
    LD = np.longdouble
    u = np.cos(np.pi * np.arange(npts) / (npts - 1))
    e = -rho * u
    if rho >= 0.02:
        # basis in u for conditioning: t = sum c_i u^i
        Phi = np.sqrt(1.0 - e)[:, None] * (u[:, None] ** np.arange(deg + 1)[None, :])
        c, _ = _lp_minimax(Phi, np.ones(npts))
        ce = c * ((-1.0 / rho) ** np.arange(deg + 1))
    else:
        # truncated Chebyshev series of (1 - e)^(-1/2) on [-rho, rho] (near-minimax; weight sqrt(x) ~ 1), in long double
        N = 64
        th = (np.arange(N, dtype=LD) + LD(0.5)) * LD(np.pi) / N
        f = (LD(1) - LD(rho) * np.cos(th)) ** LD(-0.5)
        a = [(LD(2) / N) * np.sum(f * np.cos(j * th)) for j in range(deg + 1)]
        a[0] = a[0] / 2
        # Chebyshev -> power basis in v = e / rho
        T = [np.array([1.0], dtype=LD), np.array([0.0, 1.0], dtype=LD)]
        for j in range(2, deg + 1):
            T.append(np.polynomial.polynomial.polysub(np.polynomial.polynomial.polymul(np.array([0.0, 2.0], dtype=LD), T[-1]), T[-2]))
        cv = np.zeros(deg + 1, dtype=LD)
        for j in range(deg + 1):
            cv[:len(T[j])] += a[j] * T[j]
        ce = np.array([cv[i] / LD(rho) ** i for i in range(deg + 1)], dtype=LD)
    # image in long double on a fine grid
    nfine = 4000 if rho * MARGIN <= MARGIN_TOP else 40000      # (the entries close to rho = 1: gap of 1e-5 .. 1e-3)
    uu = np.cos(np.pi * np.arange(nfine + 1) / nfine).astype(LD)
    ee = -LD(rho) * uu
    t = np.zeros_like(ee)
    for ci in ce[::-1]:
        t = t * ee + LD(ci)
    p = (LD(1) - ee) * t * t
    s, rho_out = _centre(p)
    ce = np.array([float(ci * np.sqrt(s)) for ci in ce])
    return ce, float(rho_out)


def design_start(kappa, deg, npts=1201):
    """minimax q (degree deg) for xi in [1, kappa]: max |sqrt(xi) q(xi) - 1| -> min, centred; coefficients in xi (ascending)
    and rho0"""
    if deg == 0:
        return np.array([np.sqrt(2.0 / (1.0 + kappa))]), (kappa - 1.0) / (kappa + 1.0)
    u = np.cos(np.pi * np.arange(npts) / (npts - 1))
    mid, hw = 0.5 * (kappa + 1), 0.5 * (kappa - 1)
    xi = mid + hw * u
    Phi = np.sqrt(xi)[:, None] * (u[:, None] ** np.arange(deg + 1)[None, :])
    c, _ = _lp_minimax(Phi, np.ones(npts))
    uu = np.cos(np.pi * np.arange(8001) / 8000)
    xx = mid + hw * uu
    p = xx * np.polynomial.polynomial.polyval(uu, c) ** 2
    s, rho0 = _centre(p)
    c = c * np.sqrt(s)
    P = np.polynomial.polynomial
    lin = np.array([-mid / hw, 1.0 / hw])
    out, powk = np.zeros(1), np.ones(1)
    for i in range(deg + 1):
        out = P.polyadd(out, c[i] * powk)
        powk = P.polymul(powk, lin)
    return out, float(rho0)


def build():
    # ---- rho grid: geometric in kappa' - 1 = 2 rho / (1 - rho)
    ratio = 0.88
    g = [4000.0]
    while g[-1] > 5e-7:
        g.append(g[-1] * ratio)
    top = [4000.0]                          # (extension towards rho = 1; the original entries keep their exact values)
    for _ in range(GRID_EXTRA):
        top.append(top[-1] / ratio)
    g = np.array(top[:0:-1] + g)
    rho_grid = g / (g + 2.0)               # descending
    n = len(rho_grid)
    stage = {}
    for i, rho in enumerate(rho_grid):
        for d in (1, 2, 3):
            stage[(i, d)] = design_stage(rho, d)
        if i % 20 == 0:
            print(f"  rho grid {i}/{n}", file=sys.stderr)

    def idx_of(rho):                       # smallest grid rho >= rho (grid is descending); None if above the grid
        rho = with_margin(rho)
        if rho > rho_grid[0]:
            return None
        j = int(np.searchsorted(-rho_grid, -rho, side="right")) - 1   # last index with rho_grid[j] >= rho
        return max(j, 0)

    INF = 10 ** 9
    J = [INF] * n
    act = [None] * n
    for i in range(n - 1, -1, -1):          # small rho first
        best, ba = INF, None
        for f in (1, 2, 3):
            ce, ro = stage[(i, f)]
            if ro <= TOL_RHO_OUT and f < best:
                best, ba = f, ("finish", f, ce, ro)
        for d in (1, 2, 3):
            ce, ro = stage[(i, d)]
            j = idx_of(ro)
            if j is None or j <= i:
                continue                     # no contraction on this grid
            c = d + 2 + J[j]
            if c < best:
                best, ba = c, ("stage", d, ce, ro)
        J[i], act[i] = best, ba
    # ---- start table over kappa
    kap = [1.0 + 1e-3]
    while kap[-1] < KAPPA_TABLE_MAX:
        kap.append(1.0 + (kap[-1] - 1.0) * 1.12)
    kap = np.array(kap)
    starts = []
    for kp in kap:
        best = None
        for s in (0, 1, 2):
            if s > 0 and kp < 1.05:
                continue
            a, rho0 = design_start(kp, s)
            j = idx_of(rho0)
            if j is None:
                continue
            c = 1 + s + J[j]
            if best is None or c < best[0]:
                best = (c, s, a, rho0)
        # the whole sequence of steps follows from the start (every a-priori rho does): list it, so that the kernel
        # reads its coefficients once per column instead of looking a step up after every stage
        seq, j = [], idx_of(best[3])
        while True:
            seq.append(j)
            kind, d, ce, ro = act[j]
            if kind == "finish":
                break
            j = idx_of(ro)
        assert len(seq) <= 8
        starts.append(best + (seq,))
    return rho_grid, J, act, kap, starts


def emit(path):
    rho_grid, J, act, kap, starts = build()
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_ns_schedule.py -- do not edit.  Schedule of the composite minimax polynomial\n"
                "// iteration for A^{-1/2} (see the generator's docstring and letkf_nsp.cuh).\n#pragma once\n\n")
        f.write(f"#define NSS_NRHO {len(rho_grid)}\n#define NSS_NKAPPA {len(kap)}\n")
        f.write("// per rho-grid entry (descending rho): rho, image half-width after the step, c0..c3 (in E = I - M), and\n"
                "// kind: 1..3 = stage of that degree, 11..13 = finish of degree kind - 10\n")
        f.write("struct NssStep { double rho, rho_out, c[4]; int kind, products_to_go; };\n")
        f.write("// per kappa-grid entry: start polynomial, and the indices of the steps that follow (nsteps of them, the last\n"
                "// one a finish)\n")
        f.write("struct NssStart { double kappa, rho0, a[3]; int degree, products_total; int nsteps; unsigned char step[8]; };\n")
        f.write("__constant__ NssStep nss_steps[NSS_NRHO] = {\n")
        for i, rho in enumerate(rho_grid):
            kind, d, ce, ro = act[i]
            c = list(ce) + [0.0] * (4 - len(ce))
            f.write("  {%.17g, %.17g, {%.17g, %.17g, %.17g, %.17g}, %d, %d},\n" % (rho, ro, *c, d + (10 if kind == "finish" else 0), J[i]))
        f.write("};\n__constant__ NssStart nss_starts[NSS_NKAPPA] = {\n")
        for kp, (c, s, a, rho0, seq) in zip(kap, starts):
            aa = list(a) + [0.0] * (3 - len(a))
            sq = list(seq) + [0] * (8 - len(seq))
            f.write("  {%.17g, %.17g, {%.17g, %.17g, %.17g}, %d, %d, %d, {%s}},\n"
                    % (kp, rho0, *aa, s, c, len(seq), ", ".join(str(v) for v in sq)))
        f.write("};\n")
    return rho_grid, J, act, kap, starts


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "metada_b200", "csrc", "ns_schedule_table.h")
    rho_grid, J, act, kap, starts = emit(os.path.normpath(out))
    for kp, st in zip(kap, starts):
        if any(abs(kp - t) / t < 0.07 for t in (2, 5, 10, 20, 40, 80, 160, 256)):
            print(f"kappa {kp:8.2f}: start degree {st[1]}, rho0 {st[3]:.4f}, total products {st[0]}")
