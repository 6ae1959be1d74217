"""ctypes loader for the CPU oracle (oracle/libmetada_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs -- never by the product package ``metada_b200``.

Array conventions (the reference's member-major layout):
    X      float64 [k, nz, ny, nx]  (C-contiguous)
    ox/oy/oz int32 [P]; oval, oerr float64 [P]; valid uint8 [P]
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmetada_oracle.so")

MODE_REF_COMPAT, MODE_REF_ETKF, MODE_CANONICAL = 0, 1, 2
LOC_CUTOFF, LOC_GASPARI_COHN, LOC_GAUSSIAN, LOC_EXPONENTIAL, LOC_REF_GASPARI_COHN = 0, 1, 2, 3, 4
SEM_SNAPSHOT, SEM_AS_WRITTEN = 0, 1


class LetkfParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("k", C.c_int),
        ("P", C.c_int64),
        ("radius", C.c_double), ("radius_v", C.c_double), ("inflation", C.c_double),
        ("mode", C.c_int), ("loc", C.c_int), ("use_R", C.c_int), ("semantics", C.c_int),
        ("nthreads", C.c_int), ("loc_scale", C.c_double),
    ]


class EnkfDiag(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "innovation_norm", "background_spread", "analysis_spread",
        "max_kalman_gain", "min_kalman_gain", "condition_number")]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "metada_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libmetada_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_distance_grid.restype = C.c_double
        _lib.orc_distance_grid.argtypes = [C.c_int] * 4
        _lib.orc_gaspari_cohn.restype = C.c_double
        _lib.orc_gaspari_cohn.argtypes = [C.c_double]
        _lib.orc_metrics.restype = None
        _lib.orc_loc_weight.restype = C.c_double
        _lib.orc_loc_weight.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.orc_select_local.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_distance_geo.restype = C.c_double
        _lib.orc_distance_geo.argtypes = [C.c_double] * 4
        _lib.orc_distance_cartesian.restype = C.c_double
        _lib.orc_distance_cartesian.argtypes = [C.c_double] * 6
        _lib.orc_select_local_geo.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return lib().orc_max_threads()


def gaspari_cohn(z: float) -> float:
    return lib().orc_gaspari_cohn(float(z))


def loc_weight(loc: int, dist: float, support: float, scale: float) -> float:
    return lib().orc_loc_weight(int(loc), float(dist), float(support), float(scale))


def select_local(gx, gy, ox, oy, radius):
    ox, oy = _i32(ox), _i32(oy)
    out = np.empty(len(ox), dtype=np.int32)
    c = lib().orc_select_local(C.c_int(gx), C.c_int(gy), C.c_int64(len(ox)), _p(ox, C.c_int32),
                               _p(oy, C.c_int32), C.c_double(radius), _p(out, C.c_int32))
    return out[:c].copy()


def select_counts(nx, ny, ox, oy, radius, nthreads=0):
    ox, oy = _i32(ox), _i32(oy)
    out = np.empty(nx * ny, dtype=np.int32)
    lib().orc_select_counts(C.c_int(nx), C.c_int(ny), C.c_int64(len(ox)), _p(ox, C.c_int32),
                            _p(oy, C.c_int32), C.c_double(radius), _p(out, C.c_int32),
                            C.c_int(nthreads))
    return out.reshape(ny, nx)


def hx_idw4(member, ox, oy, oz=None, valid=None):
    member = _f64(member)
    if member.ndim == 2:
        member = member[None]
    nz, ny, nx = member.shape
    ox, oy = _i32(ox), _i32(oy)
    oz = _i32(oz) if oz is not None else np.zeros(len(ox), np.int32)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    out = np.empty(len(ox))
    lib().orc_hx_idw4(_p(member, C.c_double), C.c_int(nx), C.c_int(ny), C.c_int(nz),
                      C.c_int64(len(ox)), _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32),
                      _p(v, C.c_uint8), _p(out, C.c_double))
    return out


def ensemble_mean(X):
    X = _f64(X)
    k = X.shape[0]
    n = X[0].size
    out = np.empty(X.shape[1:])
    lib().orc_ensemble_mean(_p(X, C.c_double), C.c_int(k), C.c_int64(n), _p(out, C.c_double))
    return out


def obs_space(X, ox, oy, oz, oval, valid=None):
    """Returns Y, ybar, Yp, d."""
    X = _f64(X)
    k, nz, ny, nx = X.shape
    ox, oy, oz, oval = _i32(ox), _i32(oy), _i32(oz), _f64(oval)
    P = len(ox)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    Y, Yp = np.empty((P, k)), np.empty((P, k))
    ybar, d = np.empty(P), np.empty(P)
    lib().orc_obs_space(_p(X, C.c_double), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(k),
                        C.c_int64(P), _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32),
                        _p(v, C.c_uint8), _p(oval, C.c_double), _p(Y, C.c_double),
                        _p(ybar, C.c_double), _p(Yp, C.c_double), _p(d, C.c_double))
    return Y, ybar, Yp, d


def letkf(X, ox, oy, oz, oval, oerr, valid=None, *, radius, inflation=1.0, mode=MODE_CANONICAL,
          loc=LOC_GASPARI_COHN, use_R=1, radius_v=0.0, semantics=SEM_SNAPSHOT, nthreads=0,
          cols=None, want_W=False, loc_scale=0.0):
    """Returns dict(Xa, counts[, W]).  X is not modified."""
    Xa = _f64(X).copy()
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr = _f64(oval), _f64(oerr)
    P = len(ox)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    prm = LetkfParams(nx, ny, nz, k, P, radius, radius_v, inflation, mode, loc, use_R, semantics,
                      nthreads, loc_scale)
    counts = np.full(nx * ny, -1, dtype=np.int32)
    cs = np.ascontiguousarray(cols, dtype=np.int64) if cols is not None else None
    ncols = len(cs) if cs is not None else nx * ny
    W = np.zeros((ncols, k, k)) if want_W else None
    rc = lib().orc_letkf(C.byref(prm), _p(Xa, C.c_double), _p(ox, C.c_int32), _p(oy, C.c_int32),
                         _p(oz, C.c_int32), _p(oval, C.c_double), _p(oerr, C.c_double),
                         _p(v, C.c_uint8), _p(cs, C.c_int64), C.c_int64(ncols if cs is not None else 0),
                         _p(counts, C.c_int32), _p(W, C.c_double))
    if rc:
        raise RuntimeError(f"orc_letkf failed rc={rc}")
    out = {"Xa": Xa, "counts": counts.reshape(ny, nx)}
    if want_W:
        out["W"] = W
    return out


class Ext(C.Structure):
    _fields_ = [("glat", C.POINTER(C.c_double)), ("glon", C.POINTER(C.c_double)),
                ("olat", C.POINTER(C.c_double)), ("olon", C.POINTER(C.c_double)),
                ("nvar", C.c_int), ("var_nlev", C.POINTER(C.c_int32)), ("ovar", C.POINTER(C.c_int32)),
                ("Xobs", C.POINTER(C.c_double)), ("nx_obs", C.c_int), ("ny_obs", C.c_int), ("nz_obs", C.c_int)]


def distance_cartesian(x1, y1, z1, x2, y2, z2) -> float:
    """Location::distance_to, CARTESIAN (3-D Euclid)."""
    return lib().orc_distance_cartesian(float(x1), float(y1), float(z1), float(x2), float(y2), float(z2))


def distance_geo(lat1, lon1, lat2, lon2) -> float:
    """Location::distance_to, GEOGRAPHIC (haversine km)."""
    return lib().orc_distance_geo(float(lat1), float(lon1), float(lat2), float(lon2))


def select_local_geo(clat, clon, olat, olon, radius):
    olat, olon = _f64(olat), _f64(olon)
    out = np.empty(len(olat), dtype=np.int32)
    c = lib().orc_select_local_geo(C.c_double(clat), C.c_double(clon), C.c_int64(len(olat)), _p(olat, C.c_double),
                                   _p(olon, C.c_double), C.c_double(radius), _p(out, C.c_int32), None)
    return out[:c].copy()


def select_counts_geo(glat, glon, olat, olon, radius):
    """Returns (counts [ny, nx], smallest |distance - radius| over all column/observation pairs)."""
    glat, glon, olat, olon = _f64(glat), _f64(glon), _f64(olat), _f64(olon)
    ny, nx = glat.shape
    out = np.empty(nx * ny, dtype=np.int32)
    mm = C.c_double(np.inf)
    lib().orc_select_counts_geo(C.c_int(nx), C.c_int(ny), _p(glat, C.c_double), _p(glon, C.c_double),
                                C.c_int64(len(olat)), _p(olat, C.c_double), _p(olon, C.c_double),
                                C.c_double(radius), _p(out, C.c_int32), C.byref(mm))
    return out.reshape(ny, nx), mm.value


def geo_locate(olat, olon, olev, glat, glon, vcoord=None):
    """IdentityObsOperator::convertGeographicToGrid: nearest grid point (x, y) and level z of every observation."""
    glat, glon, olat, olon = _f64(glat), _f64(glon), _f64(olat), _f64(olon)
    ny, nx = glat.shape
    P = len(olat)
    olev = _f64(olev) if olev is not None else None
    vc = _f64(vcoord) if vcoord is not None else None
    ox, oy, oz = (np.empty(P, np.int32) for _ in range(3))
    lib().orc_geo_locate(C.c_int64(P), _p(olat, C.c_double), _p(olon, C.c_double), _p(olev, C.c_double),
                         _p(glat, C.c_double), _p(glon, C.c_double), C.c_int(nx), C.c_int(ny),
                         _p(vc, C.c_double), C.c_int(len(vc) if vc is not None else 0),
                         _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32))
    return ox, oy, oz


def hx_ext(member, ox, oy, oz, var_nlev, ovar, valid=None):
    """H(x) of one member [nz_total, ny, nx] of a multi-variable state at located observation coordinates."""
    member = _f64(member)
    nz, ny, nx = member.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    vn, ov = _i32(var_nlev), _i32(ovar)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    ext = Ext(None, None, None, None, len(vn), _p(vn, C.c_int32), _p(ov, C.c_int32), None, 0, 0, 0)
    out = np.empty(len(ox))
    lib().orc_hx_ext(_p(member, C.c_double), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.byref(ext), C.c_int64(len(ox)),
                     _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32), _p(v, C.c_uint8), _p(out, C.c_double))
    return out


def letkf_ext(X, ox, oy, oz, oval, oerr, valid=None, *, radius, glat=None, glon=None, olat=None, olon=None,
              var_nlev=None, ovar=None, inflation=1.0, loc=LOC_GASPARI_COHN, use_R=1, radius_v=0.0, nthreads=0,
              loc_scale=0.0, Xobs=None, mode=MODE_CANONICAL):
    """Snapshot LETKF (canonical by default; the REF modes = LETKF.hpp:209-238 arithmetic) with GEOGRAPHIC locations (glat/glon [ny, nx], olat/olon [P]; radius in km) and / or
    a multi-variable state (var_nlev, ovar).  ox, oy, oz: nearest grid points (geo_locate).  Xobs [k, nz', ny', nx']:
    staggered grids -- H is evaluated on this ensemble (whose variables var_nlev / ovar then describe) and the
    transforms are applied to X, a single variable on its own column set glat / glon.  Returns dict(Xa, counts)."""
    Xa = _f64(X).copy()
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr = _f64(oval), _f64(oerr)
    P = len(ox)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    keep = [_f64(a) if a is not None else None for a in (glat, glon, olat, olon)]
    vn = _i32(var_nlev) if var_nlev is not None else None
    ov = _i32(ovar) if ovar is not None else None
    xo = _f64(Xobs) if Xobs is not None else None
    if xo is not None:
        assert xo.shape[0] == k
    ext = Ext(_p(keep[0], C.c_double), _p(keep[1], C.c_double), _p(keep[2], C.c_double), _p(keep[3], C.c_double),
              len(vn) if vn is not None else 0, _p(vn, C.c_int32), _p(ov, C.c_int32), _p(xo, C.c_double),
              xo.shape[3] if xo is not None else 0, xo.shape[2] if xo is not None else 0, xo.shape[1] if xo is not None else 0)
    prm = LetkfParams(nx, ny, nz, k, P, radius, radius_v, inflation, mode, loc, use_R, SEM_SNAPSHOT,
                      nthreads, loc_scale)
    counts = np.full(nx * ny, -1, dtype=np.int32)
    rc = lib().orc_letkf_ext(C.byref(prm), C.byref(ext), _p(Xa, C.c_double), _p(ox, C.c_int32), _p(oy, C.c_int32),
                             _p(oz, C.c_int32), _p(oval, C.c_double), _p(oerr, C.c_double), _p(v, C.c_uint8),
                             None, C.c_int64(0), _p(counts, C.c_int32), None)
    if rc:
        raise RuntimeError(f"orc_letkf_ext failed rc={rc}")
    return {"Xa": Xa, "counts": counts.reshape(ny, nx)}


def etkf(X, ox, oy, oz, oval, oerr, valid=None, *, inflation=1.0):
    Xa = _f64(X).copy()
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr = _f64(oval), _f64(oerr)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    rc = lib().orc_etkf(_p(Xa, C.c_double), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(k),
                        C.c_int64(len(ox)), _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32),
                        _p(oval, C.c_double), _p(oerr, C.c_double), _p(v, C.c_uint8),
                        C.c_double(inflation))
    if rc:
        raise RuntimeError(f"orc_etkf failed rc={rc}")
    return Xa


def enkf(X, ox, oy, oz, oval, oerr, Z, valid=None, *, inflation=1.0, want_gain_stats=False,
         nthreads=0):
    Xa = _f64(X).copy()
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr, Z = _f64(oval), _f64(oerr), _f64(Z)
    assert Z.shape == (len(ox), k)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    diag = EnkfDiag()
    rc = lib().orc_enkf(_p(Xa, C.c_double), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(k),
                        C.c_int64(len(ox)), _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32),
                        _p(oval, C.c_double), _p(oerr, C.c_double), _p(v, C.c_uint8),
                        C.c_double(inflation), _p(Z, C.c_double), C.c_int(int(want_gain_stats)),
                        C.c_int(nthreads), C.byref(diag))
    if rc:
        raise RuntimeError(f"orc_enkf failed rc={rc}")
    return Xa, {n: getattr(diag, n) for n, _ in EnkfDiag._fields_}


LW_UNIFORM, LW_ADAPTIVE, LW_INVERSE_VAR, LW_LIKELIHOOD = 0, 1, 2, 3
LW_WEIGHTING = {"uniform": 0, "adaptive": 1, "inverse_var": 2, "likelihood": 3}
LW_LOCFN = {"gaussian": 2, "exponential": 3, "cutoff": 0, "gaspari_cohn": 4}     # LOC_* codes


def lwenkf(X, ox, oy, oz, oval, oerr, Z, valid=None, *, inflation=1.0, radius=1.0, loc_fn=2, weighting=0):
    """LWEnKF<Tag>::Analyse (LWEnKF.hpp:207-334).  Returns (Xa, diag[9])."""
    Xa = _f64(X).copy()
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr, Z = _f64(oval), _f64(oerr), _f64(Z)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    diag = np.zeros(9)
    f = lib().orc_lwenkf
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int,
                  C.c_int, C.c_void_p, C.c_void_p]
    rc = f(Xa.ctypes.data, nx, ny, nz, k, len(ox), ox.ctypes.data, oy.ctypes.data, oz.ctypes.data, oval.ctypes.data,
           oerr.ctypes.data, v.ctypes.data if v is not None else None, inflation, radius, loc_fn, weighting, Z.ctypes.data,
           diag.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_lwenkf failed rc={rc}")
    return Xa, diag


def metrics(X, truth):
    """Metrics<double>::CalculateAll (Metrics.hpp:74-103).  X: [k, ...state], truth: [...state]."""
    X = _f64(X)
    k = X.shape[0]
    Xf = np.ascontiguousarray(X.reshape(k, -1))
    t = _f64(truth).reshape(-1)
    n = Xf.shape[1]
    mean, spread, out = np.empty(n), np.empty(n), np.empty(5)
    lib().orc_metrics(_p(Xf, C.c_double), _p(t, C.c_double), C.c_int64(n), C.c_int(k), _p(mean, C.c_double),
                      _p(spread, C.c_double), _p(out, C.c_double))
    return {"mean": mean.reshape(X.shape[1:]), "spread": spread.reshape(X.shape[1:]), "rmse": out[0], "bias": out[1],
            "correlation": out[2], "crps": out[3], "avg_spread": out[4]}


def jacobi_eigh(A):
    A = _f64(A)
    k = A.shape[0]
    ev, V = np.empty(k), np.empty((k, k))
    sw = C.c_int(0)
    rc = lib().orc_jacobi_eigh(C.c_int(k), _p(A, C.c_double), _p(ev, C.c_double), _p(V, C.c_double),
                               C.byref(sw))
    if rc:
        raise RuntimeError("jacobi failed")
    return ev, V, sw.value


def lu_inverse(A):
    A = _f64(A)
    out = np.empty_like(A)
    if lib().orc_lu_inverse(C.c_int(A.shape[0]), _p(A, C.c_double), _p(out, C.c_double)):
        raise RuntimeError("singular")
    return out


def cholesky_lower(A):
    A = _f64(A)
    out = np.empty_like(A)
    if lib().orc_cholesky_lower(C.c_int(A.shape[0]), _p(A, C.c_double), _p(out, C.c_double)):
        raise RuntimeError("not SPD")
    return out


# ---- the CPU arm of bench.py (oracle/cpu_baseline.c): built -O3 -march=native ON THE MACHINE THAT RUNS IT (the
# file name carries a hash of the CPU flags, so a library built in the build container is not reused on the GPU box)
class CpubParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("k", C.c_int), ("P", C.c_int64),
                ("radius", C.c_double), ("inflation", C.c_double), ("nthreads", C.c_int)]


_cpub = None


def cpu_baseline_lib() -> C.CDLL:
    global _cpub
    if _cpub is None:
        import hashlib
        try:
            flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
        except Exception:  # noqa: BLE001
            flags = "unknown"
        tag = hashlib.sha1(flags.encode()).hexdigest()[:10]
        src = os.path.join(_HERE, "cpu_baseline.c")
        out = os.path.join(_HERE, f"_cpub_{tag}.so")
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            cc = "gcc"
            subprocess.check_call([cc, "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", out, src, "-lm"])
        _cpub = C.CDLL(out)
        _cpub.cpub_max_threads.restype = C.c_int
    return _cpub


def cpu_baseline_letkf(X, ox, oy, oz, oval, oerr, valid=None, *, radius, inflation=1.0, nthreads=0, cols=None,
                       inplace=False):
    """Canonical LETKF (Gaspari-Cohn, symmetric square root) by the performance-minded host code.  Returns
    (Xa, sum of local observation counts).  X is not modified unless inplace (a C-contiguous float64 array)."""
    Xa = X if inplace else _f64(X).copy()
    assert Xa.dtype == np.float64 and Xa.flags.c_contiguous
    k, nz, ny, nx = Xa.shape
    ox, oy, oz = _i32(ox), _i32(oy), _i32(oz)
    oval, oerr = _f64(oval), _f64(oerr)
    v = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    prm = CpubParams(nx, ny, nz, k, len(ox), radius, inflation, nthreads)
    cs = np.ascontiguousarray(cols, dtype=np.int64) if cols is not None else None
    tot = C.c_int64(0)
    rc = cpu_baseline_lib().cpub_letkf(C.byref(prm), _p(Xa, C.c_double), _p(ox, C.c_int32), _p(oy, C.c_int32), _p(oz, C.c_int32),
                                       _p(oval, C.c_double), _p(oerr, C.c_double), _p(v, C.c_uint8), _p(cs, C.c_int64),
                                       C.c_int64(len(cs) if cs is not None else 0), C.byref(tot))
    if rc:
        raise RuntimeError(f"cpub_letkf failed rc={rc}")
    return Xa, int(tot.value)
