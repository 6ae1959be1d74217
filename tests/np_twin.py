"""NumPy twin of the oracle's local transform (numpy.linalg inv / cholesky / eigh), used only to
cross-check oracle/metada_oracle.c.  Formulas: LETKF.hpp:209-238, ETKF.hpp:150-169, Hunt 2007."""
import numpy as np


def gaspari_cohn(z):
    z = abs(z)
    if z >= 2:
        return 0.0
    if z <= 1:
        return (((-0.25 * z + 0.5) * z + 0.625) * z - 5.0 / 3.0) * z * z + 1.0
    return max(((((z / 12.0 - 0.5) * z + 0.625) * z + 5.0 / 3.0) * z - 5.0) * z + 4.0 - 2.0 / (3.0 * z), 0.0)


def hx_idw4_2d(member, ox, oy):
    ny, nx = member.shape
    out = np.empty(len(ox))
    for n, (x, y) in enumerate(zip(ox, oy)):
        x = max(0.0, min(float(nx - 1), float(x)))
        y = max(0.0, min(float(ny - 1), float(y)))
        i0, j0 = int(np.floor(x)), int(np.floor(y))
        i1, j1 = min(i0 + 1, nx - 1), min(j0 + 1, ny - 1)
        ws = wsum = 0.0
        for ii, jj in ((i0, j0), (i1, j0), (i0, j1), (i1, j1)):
            d = np.sqrt((x - ii) * (x - ii) + (y - jj) * (y - jj))
            w = 1e12 if d == 0.0 else 1.0 / d
            ws += w * member[jj, ii]
            wsum += w
        out[n] = ws / wsum
    return out


def letkf_snapshot(X, ox, oy, oval, oerr, radius, inflation, mode, loc=1, use_R=1):
    """X [k, 1, ny, nx]; modes 0 REF_COMPAT, 1 REF_ETKF, 2 CANONICAL.  2-D only."""
    k, nz, ny, nx = X.shape
    assert nz == 1
    Y = np.stack([hx_idw4_2d(X[m, 0], ox, oy) for m in range(k)], axis=1)
    ybar = Y.sum(1) / k
    Yp = Y - ybar[:, None]
    d = oval - ybar
    Xa = X.copy()
    km1 = k - 1
    for gy in range(ny):
        for gx in range(nx):
            dist = np.sqrt(((gx - ox).astype(float)) ** 2 + ((gy - oy).astype(float)) ** 2)
            idx = np.nonzero(dist <= radius)[0]
            x = X[:, 0, gy, gx]
            m = x.sum() / k
            xp = x - m
            if len(idx) == 0:
                f = inflation if mode == 1 else np.sqrt(inflation)
                Xa[:, 0, gy, gx] = m + xp * f
                continue
            Yl, dl = Yp[idx], d[idx]
            if mode == 0:
                A = Yl.T @ Yl + km1 * np.eye(k)
                Pa = np.linalg.inv(A) * inflation
                wa = Pa @ Yl.T @ dl
                Wa = np.sqrt(km1) * np.linalg.cholesky(Pa)
                Xa[:, 0, gy, gx] = m + xp * wa + xp * Wa.sum(1)
                continue
            if mode == 1:
                rinv = 1.0 / oerr[idx] ** 2
                A = (Yl.T * rinv) @ Yl + km1 * np.eye(k)
                Pa = np.linalg.inv(A)
                wa = Pa @ (Yl.T * rinv) @ dl
                Wa = np.sqrt(km1) * np.linalg.cholesky(Pa)
                xpi = xp * inflation
                Xa[:, 0, gy, gx] = m + xpi @ wa + xpi @ Wa
                continue
            rho = np.array([gaspari_cohn(dd / (0.5 * radius)) for dd in dist[idx]]) if loc == 1 else np.ones(len(idx))
            rinv = rho / (oerr[idx] ** 2 if use_R else 1.0)
            A = (Yl.T * rinv) @ Yl + (km1 / inflation) * np.eye(k)
            ev, V = np.linalg.eigh(A)
            wa = V @ ((V.T @ ((Yl.T * rinv) @ dl)) / ev)
            Wa = (V * np.sqrt(km1 / ev)) @ V.T
            Xa[:, 0, gy, gx] = m + xp @ wa + xp @ Wa
    return Xa
