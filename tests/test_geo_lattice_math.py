"""The conservative-reach argument of the geographic bucket index (metada_b200/csrc/geo_api.inl::geo_prepare_index,
geo_kernels.cuh::geo_lattice_coords), restated in NumPy and checked against the haversine distance of the oracle on
random regional domains: every observation within `radius` km of a column must (i) not be parked on the far strip
and (ii) sit within GEO_SUB + 2 lattice units of that column on both axes -- otherwise the cell walk would miss it."""
import numpy as np
import pytest

from metada_b200 import synthetic as syn
from oracle import orc

GEO_SUB = 8
R = GEO_SUB + 2


def lattice(lat, lon, radius):
    """geo_prepare_index: returns the quantiser and its bounds, or None where the library answers MDC_ERR_UNSUPPORTED."""
    rad = np.pi / 180.0
    lon_c = np.degrees(np.arctan2(np.sin(lon * rad).sum(), np.cos(lon * rad).sum()))
    u = (lon - lon_c) - 360.0 * np.rint((lon - lon_c) / 360.0)
    umin, umax, latmin, latmax = u.min(), u.max(), lat.min(), lat.max()
    delta = max(radius, 0.0) / 6371.0
    phic = max(abs(latmin), abs(latmax)) * rad
    if not (phic + delta < 0.5 * np.pi * (1.0 - 1e-6)):
        return None
    dlat = np.degrees(delta) * (1.0 + 1e-9) + 1e-12
    dlon = np.degrees(np.arcsin(min(1.0, np.sin(delta) / np.cos(phic)))) * (1.0 + 1e-9) + 1e-12
    if not (umax + dlon < 180.0 and umin - dlon > -180.0):
        return None
    ext_x, ext_y = umax - umin, latmax - latmin
    qx, qy = max(dlon / GEO_SUB, ext_x / 8192.0, 1e-9), max(dlat / GEO_SUB, ext_y / 8192.0, 1e-9)
    lo, hi_x, hi_y = -(R + 1.0), np.floor(ext_x / qx) + R + 2.0, np.floor(ext_y / qy) + R + 2.0

    def quantise(la, lo_):
        uu = (lo_ - lon_c) - 360.0 * np.rint((lo_ - lon_c) / 360.0)
        fx, fy = np.floor((uu - umin) / qx), np.floor((la - latmin) / qy)
        far = (fx < lo) | (fx > hi_x) | (fy < lo) | (fy > hi_y)
        return fx, fy, far
    return quantise


@pytest.mark.parametrize("lat0,lon0,dlat,dlon,radius", [
    (32.0, -104.0, 0.09, 0.11, 55.0),       # mid-latitude, the GPU tests' geometry
    (-48.0, 178.6, 0.09, 0.11, 60.0),       # across the dateline, southern hemisphere
    (62.0, 10.0, 0.2, 0.5, 400.0),          # high latitude, large radius: longitude reach >> latitude reach
    (-5.0, -60.0, 0.5, 0.5, 12.0),          # equator, radius far below the grid spacing
    (70.0, -150.0, 0.1, 0.4, 150.0),        # 70-76 N
    (0.0, 0.0, 0.01, 0.01, 0.0),            # radius 0
])
def test_observations_within_the_radius_are_within_the_lattice_reach(lat0, lon0, dlat, dlon, radius):
    nx, ny = 40, 30
    lat, lon = syn.geography(nx, ny, lat0=lat0, lon0=lon0, dlat=dlat, dlon=dlon)
    q = lattice(lat, lon, radius)
    assert q is not None
    rng = np.random.default_rng(int(abs(lat0) * 10 + radius))
    cfx, cfy, cfar = q(lat.ravel(), lon.ravel())
    assert not cfar.any() and cfx.min() >= -1 and cfy.min() >= -1
    # observations concentrated around the cutoff distance of random columns (the critical ones), plus a uniform cloud
    P = 4000
    cols = rng.integers(0, nx * ny, P)
    bearing = rng.uniform(0, 2 * np.pi, P)
    dist = np.where(rng.random(P) < 0.7, radius * rng.uniform(0.97, 1.03, P), radius * rng.uniform(0, 3, P))
    la1, lo1, ang = np.radians(lat.ravel()[cols]), np.radians(lon.ravel()[cols]), dist / 6371.0
    la2 = np.arcsin(np.sin(la1) * np.cos(ang) + np.cos(la1) * np.sin(ang) * np.cos(bearing))
    lo2 = lo1 + np.arctan2(np.sin(bearing) * np.sin(ang) * np.cos(la1), np.cos(ang) - np.sin(la1) * np.sin(la2))
    olat, olon = np.degrees(la2), (np.degrees(lo2) + 180.0) % 360.0 - 180.0
    ofx, ofy, ofar = q(olat, olon)
    checked = 0
    for i in range(P):
        c = cols[i]
        for cc in (c, (c + 1) % (nx * ny), (c + nx) % (nx * ny)):       # the generating column and two neighbours
            d = orc.distance_geo(lat.ravel()[cc], lon.ravel()[cc], olat[i], olon[i])
            if d <= radius:
                checked += 1
                assert not ofar[i], (i, d)
                assert abs(ofx[i] - cfx[cc]) <= R and abs(ofy[i] - cfy[cc]) <= R, (i, cc, d, ofx[i] - cfx[cc], ofy[i] - cfy[cc])
    assert checked > (300 if radius > 0 else 0)


def test_unsupported_geometries_are_recognised():
    lat, lon = syn.geography(40, 30, lat0=80.0, dlat=0.3)          # reaches 88.7 N
    assert lattice(lat, lon, 200.0) is None                        # the circle reaches the pole
    lat, lon = syn.geography(360, 10, lat0=0.0, lon0=-180.0, dlon=1.0, curvilinear=False, wrap=True)
    assert lattice(lat, lon, 100.0) is None                        # the domain wraps the longitude circle
