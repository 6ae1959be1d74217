// Internal structures of the sm_100a METADA analysis backend (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/metada_cuda_c_api.h"

struct mdc_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // public stopwatch
  cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};  // per-phase timers
  int64_t launches = 0;
  char err[512] = {0};
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  int* d_flags = nullptr;      // [16] device scratch for error flags / counters
  long long* d_stats = nullptr;  // [16]
  // second stream + events of the double-buffered row-slab transfers (copies overlap the member transposes)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_full[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_start = nullptr;
  // small results (flags, counters, statistics) come back through mapped pinned memory written by a tiny kernel:
  // a cudaMemcpy D2H of a few bytes queues on the copy engine behind the streamed pipeline's 190 MB member batches
  void* h_small = nullptr;       // host pointer, 256 bytes
  void* d_small = nullptr;       // its device alias
  long long* redo_items = nullptr;  // transforms handed from the packed Newton-Schulz kernel to its fallback
  size_t redo_cap = 0;
};

struct mdc_ens {
  mdc_ctx* ctx = nullptr;
  int nx = 0, ny = 0, nz = 0, k = 0;           // local grid (incl. halo)
  int ny_cap = 0;                              // rows allocated (mdc_ens_set_rows)
  int gx0 = 0, gy0 = 0, gnx = 0, gny = 0;      // placement in the global grid
  int own_nx = 0, own_ny = 0;                  // analysed columns: x < own_nx, y < own_ny
  double* X = nullptr;                         // [col][lev][member]
  double* mean = nullptr;                      // [col][lev] (lazily allocated)
  double* stage = nullptr;                     // staging for member-major host transfers
  size_t stage_elems = 0;
  double* host_pinned = nullptr;               // pinned bounce buffer for pageable callers
  // geography (mdc_ens_set_geography): column coordinates in degrees and the geometry's vertical coordinate
  bool geo = false;
  double *glat = nullptr, *glon = nullptr;     // [ny][nx]
  double* vcoord = nullptr;                    // [nvcoord] or nullptr
  int nvcoord = 0;
  double geo_lon_c = 0.0, geo_umin = 0.0, geo_umax = 0.0, geo_latmin = 0.0, geo_latmax = 0.0;  // host-side extents
  // grid points bucketed in the (longitude, latitude) plane for the nearest-grid-point search (geo_kernels.cuh)
  // ([0] fine cells, [1] coarse cells)
  double gc_lon0 = 0.0, gc_lat0 = 0.0, gc_lon1 = 0.0, gc_lat1 = 0.0, gc_c[2] = {0.0, 0.0};
  int gc_ncx[2] = {0, 0}, gc_ncy[2] = {0, 0};
  int32_t *gc_start[2] = {nullptr, nullptr}, *gc_pts[2] = {nullptr, nullptr};
  double *gc_plat[2] = {nullptr, nullptr}, *gc_plon[2] = {nullptr, nullptr};
  // variables (mdc_ens_set_variables): nz = sum of var_nlev; nzg = levels of the geometry (largest variable)
  int nvar = 0, nzg = 0;
  int var_off[16] = {0}, var_nlev[16] = {0};
  int32_t* levmap = nullptr;                   // [nz] level inside its variable
};

struct mdc_obs {
  mdc_ctx* ctx = nullptr;
  int64_t P = 0;       // rows in use (own + halo)
  int64_t P_own = 0;
  int64_t cap = 0;
  int k = 0;           // members of Y (set by hx / append)
  int32_t *x = nullptr, *y = nullptr, *z = nullptr;
  int64_t* gid = nullptr;
  double *val = nullptr, *err = nullptr;
  uint8_t* valid = nullptr;
  double *Y = nullptr, *ybar = nullptr, *Yp = nullptr, *d = nullptr;  // Y,Yp: [P][k]
  bool have_hx = false;
  // bucket index
  bool index_valid = false;
  int cell = 0, ncx = 0, ncy = 0, xmin = 0, ymin = 0;
  int64_t index_P = 0;
  int32_t* cell_start = nullptr;   // [ncx*ncy + 1]
  int32_t* cell_fill = nullptr;    // [ncx*ncy]
  size_t cell_cap = 0;
  int32_t* sorted_row = nullptr;   // [P] sorted position -> obs row
  int32_t *sx = nullptr, *sy = nullptr, *sz = nullptr;  // coordinates in sorted order
  int32_t* key = nullptr;          // [P]
  size_t sorted_cap = 0;
  // GEOGRAPHIC observations (mdc_obs_create_geographic): true coordinates; x, y, z above hold the nearest grid
  // point once located (mdc_obs_locate).  The index then runs on the lattice coordinates qx, qy (geo_kernels.cuh).
  bool geo = false, located = false;
  double *lat = nullptr, *lon = nullptr, *lev = nullptr;   // [P]
  int32_t* var = nullptr;                                  // [P] state variable observed, or nullptr (variable 0)
  int var_max = 0;                                         // largest entry of var (host copy, checked against the ensemble)
  int32_t *qx = nullptr, *qy = nullptr;                    // [P]
  size_t geo_cap = 0, var_cap = 0;                         // rows allocated for lat / lon / lev / qx / qy, and for var
  int32_t *cqx = nullptr, *cqy = nullptr;                  // [columns of the ensemble the index was built for]
  size_t cq_cap = 0;
  double *slat = nullptr, *slon = nullptr;                 // [P] coordinates in index order
  size_t sgeo_cap = 0;
  int geo_reach = 0;
  double geo_radius = -1.0;
  const mdc_ens* geo_ens = nullptr;
};

#define MDC_FAIL(ctx, code, ...)                              \
  do {                                                        \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
    return (code);                                            \
  } while (0)

#define MDC_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s failed at %s:%d: %s", #call, __FILE__, \
               __LINE__, cudaGetErrorString(e_));                                        \
      return MDC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MDC_LAUNCH_CHECK(ctx)                  \
  do {                                         \
    (ctx)->launches++;                         \
    MDC_CUDA(ctx, cudaGetLastError());         \
  } while (0)

static inline int mdc_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- synthetic data: pure integer hash + exact FP64 ops (bit-identical on host/numpy) ----
__host__ __device__ inline uint64_t mdc_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t mdc_hash(uint64_t seed, uint64_t idx) {
  return mdc_splitmix64(seed * 0xD1342543DE82EF95ull + idx);
}
