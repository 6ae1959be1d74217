"""NumPy emulation of the packed Newton-Schulz kernel's iteration (metada_b200/csrc/letkf_nsp.cuh,
nsp_inverse_sqrt): the same schedule tables (parsed from the generated header), the same look-up rules, the
same products -- only the upper-triangular 8 x 8 tiles of every product are formed, the lower ones implied by
symmetry.  Used by tests/test_ns_schedule.py; also documents the accuracy claims made in the kernel's comments."""
import os
import re

import numpy as np

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metada_b200", "csrc", "ns_schedule_table.h")
KAPPA_MAX = 1e5          # NSP_KAPPA_MAX_DEFAULT of the kernel


def with_margin(rho):
    """safety margin on an a-priori rho (tools/gen_ns_schedule.py: with_margin)"""
    m = rho * 1.002
    return m if m <= 4000.0 / 4002.0 else 1.0 - (1.0 - rho) * 0.90


def load_tables(path=HEADER):
    txt = open(path).read()
    steps_txt = txt[txt.index("nss_steps[NSS_NRHO] = {"):txt.index("};\n", txt.index("nss_steps[NSS_NRHO] = {"))]
    starts_txt = txt[txt.index("nss_starts[NSS_NKAPPA] = {"):]
    num = r"[-+0-9.eE]+"
    steps = [(float(v[0]), float(v[1]), [float(x) for x in v[2:6]], int(v[6]), int(v[7]))
             for v in re.findall(r"\{(%s), (%s), \{(%s), (%s), (%s), (%s)\}, (\d+), (\d+)\}" % ((num,) * 6), steps_txt)]
    starts = [(float(v[0]), float(v[1]), [float(x) for x in v[2:5]], int(v[5]), int(v[6]), [int(x) for x in v[8].split(",")][:int(v[7])])
              for v in re.findall(r"\{(%s), (%s), \{(%s), (%s), (%s)\}, (\d+), (\d+), (\d+), \{([0-9, ]+)\}\}" % ((num,) * 5), starts_txt)]
    return steps, starts


def tile_sym_product(P, Q):
    """upper-triangular tiles of the true product, lower tiles by symmetry (what the kernel stores)"""
    R = P @ Q
    nt = R.shape[0] // 8
    for I in range(nt):
        for J in range(I):
            R[8 * I:8 * I + 8, 8 * J:8 * J + 8] = R[8 * J:8 * J + 8, 8 * I:8 * I + 8].T
    return R


def start_index(starts, kappa):
    kap = np.array([s[0] for s in starts])
    i = int(np.searchsorted(kap, kappa))
    return i if i < len(kap) else -1


def step_index(steps, rho):
    rg = np.array([s[0] for s in steps])
    if not rho <= rg[0]:
        return -1
    return int(np.searchsorted(-rg, -rho, side="right")) - 1


def inverse_sqrt(A, shift, tables=None, product=tile_sym_product, kappa_max=KAPPA_MAX):
    """returns (Z, products, trace) or (None, products, reason)"""
    steps, starts = tables or load_tables()
    n = A.shape[0]
    I = np.eye(n)
    fro = np.linalg.norm(A - shift * I, "fro")
    nprod = 1
    A2 = product(A, A)
    C2 = A2 - 2 * shift * A + shift * shift * I
    hi = shift + fro
    kappa = max(min(hi, shift + np.sqrt(np.linalg.norm(C2, "fro") + 1e-13 * hi * hi)) / shift, 1.0) * (1 + 1e-9)
    si = start_index(starts, kappa) if kappa <= kappa_max else -1
    if si < 0:
        return None, nprod, "kappa"
    _, rho0, a, sdeg, _, seq = starts[si]
    rs = 1.0 / np.sqrt(shift)
    z0, z1, z2 = a[0] * rs, a[1] * rs / shift, a[2] * rs / shift ** 2
    Z = z0 * I + z1 * A + z2 * A2
    if sdeg == 0:
        E = I - z0 * z0 * A
    else:
        if sdeg == 1:
            Y = z0 * A + z1 * A2
        else:
            Y = product(A, Z); nprod += 1
        E = I - product(Z, Y); nprod += 1
    trace = [("start", sdeg, kappa, rho0)]
    # the steps listed by the start entry, in order (the kernel copies their coefficients to shared memory once)
    for j in seq:
        rg, rout, c, kind, _ = steps[j]
        d = kind % 10
        if kind > 10:
            r = np.linalg.norm(E, "fro")
            if not r * r <= n * (rg * 1.01) ** 2:
                return None, nprod, "residual"
        T = c[0] * I + c[1] * E
        if d >= 2:
            E2 = product(E, E); nprod += 1
            T = T + c[2] * E2
        if d == 3:
            T = T + c[3] * product(E, E2); nprod += 1
        Z = product(Z, T); nprod += 1
        trace.append((kind, rg, float(np.linalg.norm(E, 2))))
        if kind > 10:
            return Z, nprod, trace
        T2 = product(T, T); nprod += 1
        E = (I - T2) + product(E, T2); nprod += 1
    return None, nprod, "sequence"
