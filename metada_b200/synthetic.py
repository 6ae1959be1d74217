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


# ---- geography (the WRF-shaped case): curvilinear latitude / longitude arrays and geographic observations ----
def geography(nx, ny, lat0=32.0, lon0=-104.0, dlat=0.09, dlon=0.11, curvilinear=True, wrap=True):
    """Column coordinates in degrees, [ny, nx]: a regular latitude / longitude grid, optionally bent the way a
    projected (Lambert) WRF grid is.  wrap: longitudes folded into [-180, 180)."""
    j, i = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing="ij")
    lat = lat0 + dlat * j
    lon = lon0 + dlon * i
    if curvilinear:
        lat = lat + 0.004 * np.sin(i / 7.0)
        lon = lon0 + dlon * i * (1.0 + 0.002 * j) + 0.003 * np.cos(j / 5.0)
    if wrap:
        lon = (lon + 180.0) % 360.0 - 180.0
    return lat, lon


def geo_observations(P, lat, lon, vertical_coords=None, seed=42, sigma=0.1, margin=0.08, wrap=True):
    """P observations with GEOGRAPHIC locations, uniform over the grid's bounding box grown by `margin` of its
    size (some fall outside the domain), levels uniform over the vertical coordinate's range.  Values are drawn
    around 0 -- the tests only need seeded, well-scaled numbers."""
    rng = np.random.default_rng(seed)
    u = np.unwrap(np.deg2rad(lon), axis=1)
    u = np.rad2deg(np.unwrap(u, axis=0))
    la0, la1, lo0, lo1 = lat.min(), lat.max(), u.min(), u.max()
    dla, dlo = (la1 - la0) * margin, (lo1 - lo0) * margin
    olat = rng.uniform(la0 - dla, la1 + dla, P)
    olon = rng.uniform(lo0 - dlo, lo1 + dlo, P)
    if wrap:
        olon = (olon + 180.0) % 360.0 - 180.0
    if vertical_coords is not None:
        vc = np.asarray(vertical_coords, dtype=np.float64)
        lev = rng.uniform(vc.min() - 0.1 * np.ptp(vc), vc.max() + 0.1 * np.ptp(vc), P)
    else:
        lev = np.zeros(P)
    return {"lat": olat, "lon": olon, "level": lev, "value": sigma * 5.0 * rng.standard_normal(P),
            "err": np.full(P, sigma), "valid": np.ones(P, dtype=np.uint8)}
