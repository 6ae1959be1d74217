"""Seeded synthetic ensembles / observations (SURVEY.md section 8d shapes).

Pure integer hashing + a fixed sequence of IEEE FP64 add/mul/div, so this NumPy generator and the
device generator (``mdc_ens_fill_synthetic``, csrc/ens_kernels.cuh) are bit-identical:
    x_m(i,j,l) = wave(3i/nx) * wave(2j/ny + 1/4) * (1 + 0.01 l) + 0.5 * noise(seed + m, point)
    obs: integer (x, y, z) uniform on the grid, value = truth + sigma * noise, err = sigma.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash64(seed, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + idx)


def noise(h):
    """~N(0,1)-ish (Irwin-Hall of the four 16-bit chunks), exact in FP64."""
    h = np.asarray(h, dtype=np.uint64)
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
         ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64)
    u = s.astype(np.float64) / 65536.0 - 2.0
    return u * 1.7320508075688772


def _wave(num, den):
    f = (np.asarray(num, dtype=np.int64) % np.int64(den)).astype(np.float64) / np.float64(den)
    a = f * (1.0 - f)
    b = 1.0 - 2.0 * f
    return (10.392304845413264 * a) * b


def truth(gi, gj, lev, gnx, gny):
    gi = np.asarray(gi, dtype=np.int64)
    gj = np.asarray(gj, dtype=np.int64)
    lev = np.asarray(lev, dtype=np.int64)
    wx = _wave(3 * gi, gnx)
    wy = _wave(8 * gj + gny, 4 * gny)
    vz = 1.0 + 0.01 * lev.astype(np.float64)
    return (wx * wy) * vz


def member(m, nx, ny, nz, seed=1000, gx0=0, gy0=0, gnx=None, gny=None):
    """One member, host layout [lev][y][x] (float64)."""
    gnx = nx if gnx is None else gnx
    gny = ny if gny is None else gny
    lev, gj, gi = np.meshgrid(np.arange(nz), np.arange(gy0, gy0 + ny), np.arange(gx0, gx0 + nx),
                              indexing="ij")
    t = truth(gi, gj, lev, gnx, gny)
    idx = (lev.astype(np.uint64) * np.uint64(gny) + gj.astype(np.uint64)) * np.uint64(gnx) + gi.astype(np.uint64)
    nzv = noise(hash64(np.uint64(seed) + np.uint64(m), idx))
    return t + 0.5 * nzv


def ensemble(k, nx, ny, nz, seed=1000, **kw):
    return np.stack([member(m, nx, ny, nz, seed, **kw) for m in range(k)])


def observations(P, gnx, gny, nz, seed=42, sigma=0.1, distinct=False):
    """Returns dict(x, y, z int32; value, err float64; valid uint8)."""
    a = np.arange(P, dtype=np.uint64)
    if distinct:
        if P > gnx * gny:
            raise ValueError("more distinct obs than grid points")
        # a seeded permutation of the horizontal grid points (argsort of hashes)
        order = np.argsort(hash64(seed, np.arange(gnx * gny, dtype=np.uint64)), kind="stable")[:P]
        x = (order % gnx).astype(np.int32)
        y = (order // gnx).astype(np.int32)
    else:
        x = (hash64(seed, 3 * a) % np.uint64(gnx)).astype(np.int32)
        y = (hash64(seed, 3 * a + np.uint64(1)) % np.uint64(gny)).astype(np.int32)
    z = (hash64(seed, 3 * a + np.uint64(2)) % np.uint64(nz)).astype(np.int32)
    val = truth(x, y, z, gnx, gny) + sigma * noise(hash64(seed + 1, a))
    return dict(x=x, y=y, z=z, value=val, err=np.full(P, sigma), valid=np.ones(P, np.uint8))
