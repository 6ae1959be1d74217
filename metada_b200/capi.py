"""ctypes binding of include/metada_cuda_c_api.h (the product's C ABI).

Everything here is a 1:1 wrapper; no arithmetic happens in Python.  No fallback of any kind:
``load_library`` raises if the shared library is missing, ``Context`` raises without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = os.path.join(_HERE, "libmetada_cuda.so")
_SRC = os.path.join(_HERE, "csrc", "mdc_api.cu")
_HEADER = os.path.join(_ROOT, "include", "metada_cuda_c_api.h")

MODE_REF_COMPAT, MODE_REF_ETKF, MODE_CANONICAL = 0, 1, 2
LOC_CUTOFF, LOC_GASPARI_COHN, LOC_GAUSSIAN, LOC_EXPONENTIAL, LOC_REF_GASPARI_COHN = 0, 1, 2, 3, 4

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "shared"]   # (no -split-compile: parallel ptxas made the column
# kernels' register allocation, hence their spills and speed, vary from build to build)
LINK_FLAGS = ["-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-lpthread", "-ldl"]
# translation units: the API + every kernel except the packed Newton-Schulz column kernel, whose 42 instantiations
# are spread over six units (csrc/nsp_tu.cu compiled with a tile-count range each, csrc/nsp_launch.h) built in parallel
NSP_GROUPS = [(3, 6), (7, 9), (10, 10), (11, 12), (13, 14), (15, 16)]


class MdcError(RuntimeError):
    pass


class LetkfParams(C.Structure):
    _fields_ = [("radius", C.c_double), ("radius_v", C.c_double), ("inflation", C.c_double),
                ("mode", C.c_int), ("loc", C.c_int), ("use_R", C.c_int), ("max_sweeps", C.c_int),
                ("jacobi_tol", C.c_double), ("solver", C.c_int), ("sm_reserve", C.c_int), ("loc_scale", C.c_double),
                ("kappa_max", C.c_double)]


class LetkfStats(C.Structure):
    _fields_ = [("ms_hx", C.c_float), ("ms_index", C.c_float), ("ms_columns", C.c_float),
                ("ms_total", C.c_float), ("columns", C.c_int64), ("sum_local_obs", C.c_int64),
                ("max_local_obs", C.c_int32), ("max_sweeps", C.c_int32), ("sum_sweeps", C.c_int64),
                ("numeric_failures", C.c_int32), ("redo_transforms", C.c_int32),
                ("small_transforms", C.c_int64)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Metrics(C.Structure):
    """mdc_metrics (Metrics.hpp:15-23 MetricValues scalars)."""
    _fields_ = [(n, C.c_double) for n in ("rmse", "bias", "correlation", "crps", "avg_spread")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EnkfDiag(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "innovation_norm", "background_spread", "analysis_spread",
        "max_kalman_gain", "min_kalman_gain", "condition_number")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class LwenkfDiag(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "innovation_norm", "background_spread", "analysis_spread", "max_kalman_gain", "min_kalman_gain",
        "condition_number", "max_weight", "min_weight", "weight_variance")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class StreamConfig(C.Structure):
    _fields_ = [("gnx", C.c_int), ("gny", C.c_int), ("nz", C.c_int), ("k", C.c_int), ("row0", C.c_int), ("row1", C.c_int),
                ("slab_rows", C.c_int), ("slots", C.c_int), ("sm_reserve", C.c_int), ("radius", C.c_double)]


def lib_path() -> str:
    return _LIB


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a (works without a GPU): objects in metada_b200/_obj, units in parallel."""
    from concurrent.futures import ThreadPoolExecutor
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [_HEADER]
    newest = max(os.path.getmtime(s) for s in srcs)
    objdir = os.path.join(_HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    units = [("mdc_api.o", _SRC, []), ("mdc_runtime.o", os.path.join(csrc, "mdc_runtime.cpp"), [])]
    units += [(f"nsp_{lo}_{hi}.o", os.path.join(csrc, "nsp_tu.cu"), [f"-DNSP_LO={lo}", f"-DNSP_HI={hi}"]) for lo, hi in NSP_GROUPS]
    only = os.environ.get("MDC_BUILD_ONLY")          # development: rebuild just these objects (comma-separated)

    def compile_unit(u):
        obj, src, defs = u
        out = os.path.join(objdir, obj)
        stale = force or not os.path.exists(out) or os.path.getmtime(out) < newest
        if only is not None:
            stale = obj in only.split(",") or not os.path.exists(out)
        if stale:
            cmd = ["nvcc", *NVCC_FLAGS, *defs, "-c", "-o", out, src]
            if os.environ.get("MDC_NSP_PROFILE") and obj.startswith("nsp_"):
                cmd.insert(1, "-DNSP_PROFILE")          # per-phase clock ticks (tools/nsp_phase_profile.py)
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.check_call(cmd)
        return out, stale

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_unit, units))
    if any(st for _, st in res) or not os.path.exists(_LIB):
        subprocess.check_call(["nvcc", *LINK_FLAGS, "-o", _LIB, *[o for o, _ in res]])
    return _LIB


def exported_symbols() -> list[str]:
    """extern "C" entry points declared in the public header."""
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mdc_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MDC_LIB", _LIB)     # development: an instrumented build of the same sources (tools/build_prof.sh)
    if not os.path.exists(path):
        raise MdcError(f"{path} is missing: build it with metada_b200.build_library() "
                       "(__graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    pd = C.POINTER(C.c_double)
    sig = {
        "mdc_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "mdc_ctx_destroy": (C.c_int, [vp]),
        "mdc_last_error": (C.c_char_p, [vp]),
        "mdc_ctx_sync": (C.c_int, [vp]),
        "mdc_ctx_stream": (vp, [vp]),
        "mdc_timer_start": (C.c_int, [vp]),
        "mdc_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "mdc_ctx_launch_count": (i64, [vp]),
        "mdc_ctx_sm_count": (C.c_int, [vp]),
        "mdc_ctx_last_stats": (C.c_int, [vp, C.POINTER(C.c_int64)]),
        "mdc_ctx_flush_l2": (C.c_int, [vp]),
        "mdc_ens_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "mdc_ens_destroy": (C.c_int, [vp]),
        "mdc_ens_set_domain": (C.c_int, [vp] + [C.c_int] * 6),
        "mdc_ens_upload_member": (C.c_int, [vp, C.c_int, vp]),
        "mdc_ens_download_member": (C.c_int, [vp, C.c_int, vp]),
        "mdc_ens_upload_members": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "mdc_ens_download_members": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "mdc_ens_upload_members_rows": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.c_int, C.c_int]),
        "mdc_ens_download_members_rows": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int]),
        "mdc_dev_malloc": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "mdc_dev_free": (C.c_int, [vp, vp]),
        "mdc_ens_fill_synthetic": (C.c_int, [vp, C.c_uint64]),
        "mdc_ens_mean": (C.c_int, [vp, vp]),
        "mdc_ens_checksum": (C.c_int, [vp, pd, pd]),
        "mdc_ens_metrics": (C.c_int, [vp, vp, C.POINTER(Metrics), vp]),
        "mdc_ens_devptr": (vp, [vp]),
        "mdc_ens_bytes": (i64, [vp]),
        "mdc_obs_create": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "mdc_obs_assign": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "mdc_ens_set_rows": (C.c_int, [vp, C.c_int]),
        "mdc_obs_destroy": (C.c_int, [vp]),
        "mdc_obs_size": (i64, [vp]),
        "mdc_ens_set_geography": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "mdc_ens_set_variables": (C.c_int, [vp, C.c_int, vp]),
        "mdc_obs_set_variables": (C.c_int, [vp, vp]),
        "mdc_obs_create_geographic": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "mdc_obs_locate": (C.c_int, [vp, vp]),
        "mdc_obs_download_grid_coords": (C.c_int, [vp, vp, vp, vp]),
        "mdc_hx_idw4": (C.c_int, [vp, vp]),
        "mdc_hx_download": (C.c_int, [vp, vp, vp, vp, vp]),
        "mdc_obs_pack_rows": (C.c_int, [vp, C.c_int, C.c_int, vp, i64, C.POINTER(i64)]),
        "mdc_obs_pack_rows_geo": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp, i64, C.POINTER(i64)]),
        "mdc_ens_set_geography_from": (C.c_int, [vp, vp]),
        "mdc_ens_geography_frame": (C.c_int, [vp] + [C.POINTER(C.c_double)] * 5),
        "mdc_obs_append_rows": (C.c_int, [vp, vp, i64]),
        "mdc_obs_row_doubles": (C.c_int, [vp]),
        "mdc_obs_index_build": (C.c_int, [vp, C.c_int]),
        "mdc_obs_index_query_counts": (C.c_int, [vp, vp, dbl, vp]),
        "mdc_obs_index_query_lists": (C.c_int, [vp, vp, dbl, vp, i64, C.c_int32, vp, vp]),
        "mdc_letkf_analyse": (C.c_int, [vp, vp, C.POINTER(LetkfParams), C.POINTER(LetkfStats)]),
        "mdc_letkf_column_transform": (C.c_int, [vp, vp, C.POINTER(LetkfParams), i64, vp]),
        "mdc_etkf_analyse": (C.c_int, [vp, vp, dbl]),
        "mdc_enkf_analyse": (C.c_int, [vp, vp, dbl, vp, C.c_uint64, C.c_int, C.POINTER(EnkfDiag)]),
        "mdc_dev_copy": (C.c_int, [vp, vp, vp, i64, C.c_int]),
        "mdc_stream_create": (C.c_int, [C.c_int, C.POINTER(StreamConfig), C.POINTER(vp)]),
        "mdc_stream_destroy": (C.c_int, [vp]),
        "mdc_stream_last_error": (C.c_char_p, [vp]),
        "mdc_stream_slabs": (C.c_int, [vp]),
        "mdc_stream_slots": (C.c_int, [vp]),
        "mdc_stream_analyse": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.c_int, i64, vp, vp, vp, vp, vp, vp,
                                          C.POINTER(LetkfParams), C.POINTER(LetkfStats)]),
        "mdc_stream_timings": (C.c_int, [vp, pd, pd, C.POINTER(i64)]),
        "mdc_comm_get_unique_id": (C.c_int, [vp, C.c_int]),
        "mdc_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "mdc_comm_max": (C.c_int, [vp, pd]),
        "mdc_comm_allgather_rows": (C.c_int, [vp, C.POINTER(vp)]),
        "mdc_lwenkf_analyse": (C.c_int, [vp, vp, dbl, dbl, C.c_int, C.c_int, vp, C.c_uint64, C.POINTER(LwenkfDiag)]),
        "mdc_bench_fp64_fma": (C.c_int, [vp, pd]),
        "mdc_bench_fp64_dmma": (C.c_int, [vp, pd]),
        "mdc_bench_hbm_copy": (C.c_int, [vp, pd]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Context:
    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.mdc_ctx_create(device, C.byref(h))
        if rc:
            raise MdcError(f"mdc_ctx_create(device={device}) failed rc={rc}: no usable CUDA device "
                           "(this backend has no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc:
            raise MdcError(f"rc={rc}: {self.L.mdc_last_error(self.h).decode()}")

    def sync(self):
        self.check(self.L.mdc_ctx_sync(self.h))

    def stream_ptr(self) -> int:
        return int(self.L.mdc_ctx_stream(self.h) or 0)

    def timer_start(self):
        self.check(self.L.mdc_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(self.L.mdc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.L.mdc_ctx_launch_count(self.h))

    def sm_count(self) -> int:
        return int(self.L.mdc_ctx_sm_count(self.h))

    def last_stats(self):
        out = (C.c_int64 * 16)()
        self.check(self.L.mdc_ctx_last_stats(self.h, out))
        return [int(v) for v in out]

    def flush_l2(self):
        self.check(self.L.mdc_ctx_flush_l2(self.h))

    def bench_fp64_fma(self) -> float:
        v = C.c_double()
        self.check(self.L.mdc_bench_fp64_fma(self.h, C.byref(v)))
        return v.value

    def bench_fp64_dmma(self) -> float:
        v = C.c_double()
        self.check(self.L.mdc_bench_fp64_dmma(self.h, C.byref(v)))
        return v.value

    def bench_hbm_copy(self) -> float:
        v = C.c_double()
        self.check(self.L.mdc_bench_hbm_copy(self.h, C.byref(v)))
        return v.value

    def dev_malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.L.mdc_dev_malloc(self.h, nbytes, C.byref(p)))
        return int(p.value or 0)

    def dev_free(self, ptr: int):
        self.check(self.L.mdc_dev_free(self.h, C.c_void_p(ptr)))

    def close(self):
        if self.h:
            self.L.mdc_ctx_destroy(self.h)
            self.h = None


class Ensemble:
    """Device-resident ensemble store, layout [col][lev][member]."""

    def __init__(self, ctx: Context, nx: int, ny: int, nz: int, k: int):
        self.ctx, self.nx, self.ny, self.nz, self.k = ctx, nx, ny, nz, k
        h = C.c_void_p()
        ctx.check(ctx.L.mdc_ens_create(ctx.h, nx, ny, nz, k, C.byref(h)))
        self.h = h

    def set_domain(self, gx0, gy0, gnx, gny, own_nx, own_ny):
        self.ctx.check(self.ctx.L.mdc_ens_set_domain(self.h, gx0, gy0, gnx, gny, own_nx, own_ny))

    def set_rows(self, ny: int):
        self.ctx.check(self.ctx.L.mdc_ens_set_rows(self.h, ny))
        self.ny = ny

    def upload(self, X: np.ndarray):
        """X: [k, nz, ny, nx] float64 (any host memory)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        assert X.shape == (self.k, self.nz, self.ny, self.nx), X.shape
        self.upload_ptrs(0, [X[m].ctypes.data for m in range(self.k)])
        self.ctx.sync()

    def upload_ptrs(self, m0: int, ptrs):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self.ctx.check(self.ctx.L.mdc_ens_upload_members(self.h, m0, len(ptrs), arr))

    def download_ptrs(self, m0: int, ptrs):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self.ctx.check(self.ctx.L.mdc_ens_download_members(self.h, m0, len(ptrs), arr))

    def upload_rows(self, ptrs, host_ny: int, host_y0: int, m0: int = 0):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self.ctx.check(self.ctx.L.mdc_ens_upload_members_rows(self.h, m0, len(ptrs), arr, host_ny, host_y0))

    def download_rows(self, ptrs, host_ny: int, host_y0: int, nrows: int, m0: int = 0):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self.ctx.check(self.ctx.L.mdc_ens_download_members_rows(self.h, m0, len(ptrs), arr, host_ny, host_y0, nrows))

    def download(self) -> np.ndarray:
        out = np.empty((self.k, self.nz, self.ny, self.nx))
        self.download_ptrs(0, [out[m].ctypes.data for m in range(self.k)])
        return out

    def upload_member(self, m: int, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.ctx.check(self.ctx.L.mdc_ens_upload_member(self.h, m, _ptr(x)))
        self.ctx.sync()

    def download_member(self, m: int) -> np.ndarray:
        out = np.empty((self.nz, self.ny, self.nx))
        self.ctx.check(self.ctx.L.mdc_ens_download_member(self.h, m, _ptr(out)))
        return out

    def set_geography(self, lat, lon, vertical_coords=None):
        """Column coordinates in degrees, [ny, nx] (the geometry's 2-D latitude / longitude arrays), and the
        geometry's vertical coordinate (nearest-level lookup of geographic observations)."""
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        assert lat.shape == (self.ny, self.nx) and lon.shape == (self.ny, self.nx), (lat.shape, lon.shape)
        vc = np.ascontiguousarray(vertical_coords, dtype=np.float64) if vertical_coords is not None else None
        self.ctx.check(self.ctx.L.mdc_ens_set_geography(self.h, _ptr(lat), _ptr(lon), len(vc) if vc is not None else 0, _ptr(vc)))

    def set_geography_from(self, whole: "Ensemble"):
        """Geography of a decomposed store = its window of `whole` (a store covering the whole grid), with the
        global frame: the analysis is then bit-identical to the single-store one."""
        self.ctx.check(self.ctx.L.mdc_ens_set_geography_from(self.h, whole.h))

    def geography_frame(self) -> dict:
        v = [C.c_double() for _ in range(5)]
        self.ctx.check(self.ctx.L.mdc_ens_geography_frame(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("lon_c", "umin", "umax", "latmin", "latmax"), (x.value for x in v)))

    def set_variables(self, var_nlev):
        """Variables of the state: a member is [var][lev][y][x], nz = sum(var_nlev)."""
        vn = np.ascontiguousarray(var_nlev, dtype=np.int32)
        self.ctx.check(self.ctx.L.mdc_ens_set_variables(self.h, len(vn), _ptr(vn)))

    def fill_synthetic(self, seed: int = 1000):
        self.ctx.check(self.ctx.L.mdc_ens_fill_synthetic(self.h, seed))

    def mean(self) -> np.ndarray:
        out = np.empty((self.nz, self.ny, self.nx))
        self.ctx.check(self.ctx.L.mdc_ens_mean(self.h, _ptr(out)))
        return out

    def checksum(self):
        s, s2 = C.c_double(), C.c_double()
        self.ctx.check(self.ctx.L.mdc_ens_checksum(self.h, C.byref(s), C.byref(s2)))
        return s.value, s2.value

    def metrics(self, truth: np.ndarray, want_spread: bool = False) -> dict:
        """Verification metrics of this ensemble against truth [nz, ny, nx] (Metrics.hpp:74-103)."""
        t = Ensemble(self.ctx, self.nx, self.ny, self.nz, 1)
        try:
            t.upload_member(0, np.ascontiguousarray(truth, dtype=np.float64).reshape(self.nz, self.ny, self.nx))
            m = Metrics()
            sp = np.empty((self.nz, self.ny, self.nx)) if want_spread else None
            self.ctx.check(self.ctx.L.mdc_ens_metrics(self.h, t.h, C.byref(m), _ptr(sp) if sp is not None else None))
        finally:
            t.close()
        out = m.asdict()
        if want_spread:
            out["spread"] = sp
        return out

    def devptr(self) -> int:
        return int(self.ctx.L.mdc_ens_devptr(self.h))

    def nbytes(self) -> int:
        return int(self.ctx.L.mdc_ens_bytes(self.h))

    def close(self):
        if self.h:
            self.ctx.L.mdc_ens_destroy(self.h)
            self.h = None


class Observations:
    """Device SoA of GRID observations (+ Y', d after hx)."""

    def __init__(self, ctx: Context, x, y, z, value, err, valid=None, gid=None):
        self.ctx = ctx
        x = np.ascontiguousarray(x, dtype=np.int32)
        y = np.ascontiguousarray(y, dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.int32) if z is not None else None
        value = np.ascontiguousarray(value, dtype=np.float64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        valid = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
        gid = np.ascontiguousarray(gid, dtype=np.int64) if gid is not None else None
        h = C.c_void_p()
        ctx.check(ctx.L.mdc_obs_create(ctx.h, len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(value),
                                       _ptr(err), _ptr(valid), _ptr(gid), C.byref(h)))
        self.h = h

    @classmethod
    def geographic(cls, ctx: Context, lat, lon, level, value, err, valid=None, gid=None) -> "Observations":
        """Observations with GEOGRAPHIC locations (degrees, level in the geometry's vertical coordinate); the local
        selection is then by haversine kilometres (Location.hpp:213-217)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        level = np.ascontiguousarray(level, dtype=np.float64) if level is not None else None
        value = np.ascontiguousarray(value, dtype=np.float64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        valid = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
        gid = np.ascontiguousarray(gid, dtype=np.int64) if gid is not None else None
        h = C.c_void_p()
        ctx.check(ctx.L.mdc_obs_create_geographic(ctx.h, len(lat), _ptr(lat), _ptr(lon), _ptr(level), _ptr(value),
                                                  _ptr(err), _ptr(valid), _ptr(gid), C.byref(h)))
        self.h = h
        return self

    def set_variables(self, var):
        """State variable each observation observes (index into Ensemble.set_variables); None = variable 0."""
        v = np.ascontiguousarray(var, dtype=np.int32) if var is not None else None
        assert v is None or len(v) == self.size()
        self.ctx.check(self.ctx.L.mdc_obs_set_variables(self.h, _ptr(v)))

    def locate(self, ens: "Ensemble"):
        """Nearest grid point and level of every geographic observation (IdentityObsOperator.hpp:484-530)."""
        self.ctx.check(self.ctx.L.mdc_obs_locate(self.h, ens.h))

    def grid_coords(self):
        P = self.size()
        x, y, z = (np.empty(P, np.int32) for _ in range(3))
        self.ctx.check(self.ctx.L.mdc_obs_download_grid_coords(self.h, _ptr(x), _ptr(y), _ptr(z)))
        return x, y, z

    def assign(self, x, y, z, value, err, valid=None, gid=None):
        x = np.ascontiguousarray(x, dtype=np.int32)
        y = np.ascontiguousarray(y, dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.int32) if z is not None else None
        value = np.ascontiguousarray(value, dtype=np.float64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        valid = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
        gid = np.ascontiguousarray(gid, dtype=np.int64) if gid is not None else None
        self.ctx.check(self.ctx.L.mdc_obs_assign(self.h, len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(value),
                                                 _ptr(err), _ptr(valid), _ptr(gid)))

    def size(self) -> int:
        return int(self.ctx.L.mdc_obs_size(self.h))

    def hx(self, ens: Ensemble):
        self.ctx.check(self.ctx.L.mdc_hx_idw4(ens.h, self.h))
        self._k = ens.k

    def hx_download(self, want=("Y", "ybar", "Yp", "d")):
        P, k = self.size(), self._k
        out = {}
        if "Y" in want:
            out["Y"] = np.empty((P, k))
        if "ybar" in want:
            out["ybar"] = np.empty(P)
        if "Yp" in want:
            out["Yp"] = np.empty((P, k))
        if "d" in want:
            out["d"] = np.empty(P)
        self.ctx.check(self.ctx.L.mdc_hx_download(self.h, _ptr(out.get("Y")), _ptr(out.get("ybar")),
                                                  _ptr(out.get("Yp")), _ptr(out.get("d"))))
        return out

    def row_doubles(self) -> int:
        return int(self.ctx.L.mdc_obs_row_doubles(self.h))

    def pack_rows(self, ylo: int, yhi: int, dev_ptr: int, cap: int) -> int:
        n = C.c_int64()
        self.ctx.check(self.ctx.L.mdc_obs_pack_rows(self.h, ylo, yhi, C.c_void_p(dev_ptr), cap, C.byref(n)))
        return n.value

    def pack_rows_geo(self, lat_lo, lat_hi, u_lo, u_hi, lon_c, dev_ptr: int, cap: int) -> int:
        """Own observations inside a box of the geography's frame (see mdc_obs_pack_rows_geo); returns the count found
        (rows beyond cap are dropped: call with cap = 0 to count)."""
        n = C.c_int64()
        self.ctx.check(self.ctx.L.mdc_obs_pack_rows_geo(self.h, lat_lo, lat_hi, u_lo, u_hi, lon_c, C.c_void_p(dev_ptr), cap, C.byref(n)))
        return n.value

    def append_rows(self, dev_ptr: int, n: int):
        self.ctx.check(self.ctx.L.mdc_obs_append_rows(self.h, C.c_void_p(dev_ptr), n))

    def index_build(self, cell: int):
        self.ctx.check(self.ctx.L.mdc_obs_index_build(self.h, cell))

    def query_counts(self, ens: Ensemble, radius: float) -> np.ndarray:
        out = np.empty(ens.nx * ens.ny, dtype=np.int32)
        self.ctx.check(self.ctx.L.mdc_obs_index_query_counts(self.h, ens.h, radius, _ptr(out)))
        return out.reshape(ens.ny, ens.nx)

    def query_lists(self, ens: Ensemble, radius: float, cols, cap: int):
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        lists = np.empty((len(cols), cap), dtype=np.int64)
        counts = np.empty(len(cols), dtype=np.int32)
        self.ctx.check(self.ctx.L.mdc_obs_index_query_lists(self.h, ens.h, radius, _ptr(cols), len(cols),
                                                            cap, _ptr(lists), _ptr(counts)))
        return [lists[i, :min(counts[i], cap)].copy() for i in range(len(cols))], counts

    def close(self):
        if self.h:
            self.ctx.L.mdc_obs_destroy(self.h)
            self.h = None


SOLVER_AUTO, SOLVER_JACOBI, SOLVER_NEWTON_SCHULZ, SOLVER_NEWTON_SCHULZ_FULL = 0, 1, 2, 3


def make_params(radius, inflation=1.0, mode=MODE_CANONICAL, loc=LOC_GASPARI_COHN, use_R=1,
                radius_v=0.0, max_sweeps=0, jacobi_tol=0.0, solver=SOLVER_AUTO, sm_reserve=0,
                loc_scale=0.0, kappa_max=0.0) -> LetkfParams:
    p = LetkfParams()
    p.radius, p.radius_v, p.inflation = radius, radius_v, inflation
    p.mode, p.loc, p.use_R, p.max_sweeps, p.jacobi_tol = mode, loc, use_R, max_sweeps, jacobi_tol
    p.solver = solver
    p.sm_reserve = sm_reserve
    p.loc_scale = loc_scale
    p.kappa_max = kappa_max
    return p


def letkf_analyse(ens: Ensemble, obs: Observations, params: LetkfParams) -> dict:
    st = LetkfStats()
    ens.ctx.check(ens.ctx.L.mdc_letkf_analyse(ens.h, obs.h, C.byref(params), C.byref(st)))
    return st.asdict()


def letkf_column_transform(ens: Ensemble, obs: Observations, params: LetkfParams, col: int) -> np.ndarray:
    W = np.empty((ens.k, ens.k))
    ens.ctx.check(ens.ctx.L.mdc_letkf_column_transform(ens.h, obs.h, C.byref(params), col, _ptr(W)))
    return W


def etkf_analyse(ens: Ensemble, obs: Observations, inflation: float):
    ens.ctx.check(ens.ctx.L.mdc_etkf_analyse(ens.h, obs.h, inflation))


def enkf_analyse(ens: Ensemble, obs: Observations, inflation: float, Z=None, seed: int = 7,
                 want_gain_stats: bool = False) -> dict:
    Zc = np.ascontiguousarray(Z, dtype=np.float64) if Z is not None else None
    d = EnkfDiag()
    ens.ctx.check(ens.ctx.L.mdc_enkf_analyse(ens.h, obs.h, inflation, _ptr(Zc), seed,
                                             int(want_gain_stats), C.byref(d)))
    return d.asdict()

LW_UNIFORM, LW_ADAPTIVE, LW_INVERSE_VAR, LW_LIKELIHOOD = 0, 1, 2, 3


def lwenkf_analyse(ens: Ensemble, obs: Observations, inflation: float, loc_radius: float, loc_fn: int, weighting: int,
                   Z=None, seed: int = 7) -> dict:
    Zc = np.ascontiguousarray(Z, dtype=np.float64) if Z is not None else None
    d = LwenkfDiag()
    ens.ctx.check(ens.ctx.L.mdc_lwenkf_analyse(ens.h, obs.h, inflation, loc_radius, loc_fn, weighting, _ptr(Zc), seed, C.byref(d)))
    return d.asdict()


class Stream:
    """mdc_stream: the C++ streaming / sharding runtime (csrc/mdc_runtime.cpp).  Host members in, analysed in place."""

    def __init__(self, device, gnx, gny, nz, k, radius, row_range=None, slab_rows=0, slots=4, sm_reserve=8):
        self.L = load_library()
        r0, r1 = (0, gny) if row_range is None else row_range
        self.cfg = StreamConfig(gnx, gny, nz, k, r0, r1, slab_rows, slots, sm_reserve, float(radius))
        h = C.c_void_p()
        rc = self.L.mdc_stream_create(device, C.byref(self.cfg), C.byref(h))
        if rc:
            raise MdcError(f"mdc_stream_create rc={rc}")
        self.h = h

    def check(self, rc):
        if rc:
            raise MdcError(f"rc={rc}: {self.L.mdc_stream_last_error(self.h).decode()}")

    @property
    def nslab(self):
        return self.L.mdc_stream_slabs(self.h)

    @property
    def nslots(self):
        return self.L.mdc_stream_slots(self.h)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().mdc_comm_get_unique_id(buf, 128)
        if rc:
            raise MdcError(f"mdc_comm_get_unique_id rc={rc} (NCCL missing?)")
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = C.create_string_buffer(uid, 128)
        self.check(self.L.mdc_comm_init(self.h, buf, rank, world))

    def allgather_rows(self, member_ptrs):
        arr = (C.c_void_p * len(member_ptrs))(*member_ptrs)
        self.check(self.L.mdc_comm_allgather_rows(self.h, arr))

    def comm_max(self, v: float) -> float:
        x = C.c_double(v)
        self.check(self.L.mdc_comm_max(self.h, C.byref(x)))
        return x.value

    def analyse(self, member_ptrs, obs: dict, params, host_row0=0, host_ny=None) -> dict:
        host_ny = self.cfg.gny if host_ny is None else host_ny
        arr = (C.c_void_p * len(member_ptrs))(*member_ptrs)
        x = np.ascontiguousarray(obs["x"], np.int32); y = np.ascontiguousarray(obs["y"], np.int32)
        z = np.ascontiguousarray(obs["z"], np.int32) if obs.get("z") is not None else None
        v = np.ascontiguousarray(obs["value"], np.float64); e = np.ascontiguousarray(obs["err"], np.float64)
        ok = np.ascontiguousarray(obs["valid"], np.uint8) if obs.get("valid") is not None else None
        st = LetkfStats()
        self.check(self.L.mdc_stream_analyse(self.h, arr, host_row0, host_ny, len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(v), _ptr(e),
                                             _ptr(ok), C.byref(params), C.byref(st)))
        return st.asdict()

    def timings(self) -> dict:
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self.L.mdc_stream_timings(self.h, C.byref(a), C.byref(b), C.byref(n))
        return {"edge_halo_ms": a.value, "stream_ms": b.value, "halo_rows": n.value}

    def close(self):
        if self.h:
            self.L.mdc_stream_destroy(self.h)
            self.h = None
