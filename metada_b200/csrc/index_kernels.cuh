// Device-side bucketed spatial index over the observations (grid-cell hashing, counting sort by
// cell with a scan, ids ascending inside each cell).  Replaces the O(P) linear scan per grid point
// of LETKF.hpp:159-165.  Integer/byte work, HBM-bound and tiny next to the column kernel.
//
//   key(i)     = ((y_i - ymin) / cell) * ncx + (x_i - xmin) / cell
//   cell_start = exclusive scan of the per-cell histogram
//   sorted_row = obs rows ordered by (key, gid)   -> deterministic, independent of atomics order
//
// A column query walks the cell rows overlapping [gy - r, gy + r]; inside one cell row the
// candidate cells are contiguous in the sorted order, so each cell row is ONE contiguous range.
#pragma once
#include "mdc_internal.cuh"

__global__ void index_bbox_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                  int64_t P, int* __restrict__ bbox /*xmin,ymin,xmax,ymax*/) {
  int xmin = INT_MAX, ymin = INT_MAX, xmax = INT_MIN, ymax = INT_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    xmin = min(xmin, x[i]); xmax = max(xmax, x[i]);
    ymin = min(ymin, y[i]); ymax = max(ymax, y[i]);
  }
  for (int o = 16; o; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(bbox + 0, xmin); atomicMin(bbox + 1, ymin);
    atomicMax(bbox + 2, xmax); atomicMax(bbox + 3, ymax);
  }
}

__global__ void index_key_hist_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                      int64_t P, int xmin, int ymin, int cell, int ncx,
                                      int32_t* __restrict__ key, int32_t* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    int kx = (x[i] - xmin) / cell, ky = (y[i] - ymin) / cell;
    int kk = ky * ncx + kx;
    key[i] = kk;
    atomicAdd(hist + kk, 1);
  }
}

// single-block exclusive scan (ncell is ~1e4..1e5): out[0..n] with out[n] = total
__global__ void index_scan_kernel(const int32_t* __restrict__ hist, int32_t* __restrict__ out, int n) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < n) ? hist[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int ws = (lane < nw) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_sums[lane] = ws;  // inclusive
    }
    __syncthreads();
    int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + (s - v);
    if (i < n) out[i] = prefix;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void index_scatter_kernel(const int32_t* __restrict__ key, int64_t P,
                                     const int32_t* __restrict__ cell_start,
                                     int32_t* __restrict__ fill, int32_t* __restrict__ sorted_row) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    int kk = key[i];
    int pos = cell_start[kk] + atomicAdd(fill + kk, 1);
    sorted_row[pos] = (int32_t)i;
  }
}

// order each cell's segment by global obs id (insertion sort for the usual tens of entries, heapsort beyond) and
// emit the coordinates in sorted order for sequential reads in the column query
__global__ void index_cell_sort_kernel(const int32_t* __restrict__ cell_start, int ncell,
                                       int32_t* __restrict__ sorted_row,
                                       const int64_t* __restrict__ gid,
                                       const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                       const int32_t* __restrict__ z, int32_t* __restrict__ sx,
                                       int32_t* __restrict__ sy, int32_t* __restrict__ sz) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int b = cell_start[c], e = cell_start[c + 1];
  if (e - b <= 48) {
    for (int a = b + 1; a < e; ++a) {
      int32_t r = sorted_row[a];
      int64_t gr = gid[r];
      int q = a - 1;
      while (q >= b && gid[sorted_row[q]] > gr) { sorted_row[q + 1] = sorted_row[q]; --q; }
      sorted_row[q + 1] = r;
    }
  } else {
    // a crowded cell (clustered reports: swaths, many reports at one station): heapsort, O(n log n) -- the insertion
    // sort above is O(n^2) and would take minutes on 1e5 rows
    int32_t* h = sorted_row + b;
    const int n = e - b;
    auto sift = [&](int start, int end) {
      int root = start;
      while (2 * root + 1 < end) {
        int child = 2 * root + 1;
        if (child + 1 < end && gid[h[child]] < gid[h[child + 1]]) ++child;
        if (gid[h[root]] >= gid[h[child]]) return;
        const int32_t t = h[root]; h[root] = h[child]; h[child] = t;
        root = child;
      }
    };
    for (int st = n / 2 - 1; st >= 0; --st) sift(st, n);
    for (int end = n - 1; end > 0; --end) {
      const int32_t t = h[0]; h[0] = h[end]; h[end] = t;
      sift(0, end);
    }
  }
  for (int a = b; a < e; ++a) {
    int32_t r = sorted_row[a];
    sx[a] = x[r]; sy[a] = y[r]; sz[a] = z[r];
  }
}

struct IndexView {
  const int32_t* cell_start;
  const int32_t* sorted_row;
  const int32_t *sx, *sy, *sz;
  int cell, ncx, ncy, xmin, ymin;
  // GEOGRAPHIC observations (geo_kernels.cuh): the index runs on a quantised lat/lon lattice (sx, sy = lattice
  // coordinates of the observations, cqx, cqy = of the columns, reach = conservative integer reach of the selection
  // radius in lattice units) and the selection test is the haversine distance of the true coordinates.
  int geo, reach;
  const int32_t *cqx, *cqy;      // [local column]
  const double *clat, *clon;     // [local column] degrees
  const double *slat, *slon;     // [sorted position] degrees
  const int32_t* levmap;         // [total level] -> level inside its variable (multi-variable states), or nullptr
};

// integer reach of the selection radius in index units, and the column's coordinates in the index
// EXT = false is the GRID / single-variable path, compiled exactly as before these fields existed (the column
// kernels' register allocation is fragile); EXT = true instantiations serve geographic and multi-variable analyses.
template <bool EXT>
__device__ __forceinline__ int index_reach(const IndexView& iv, double radius) {
  if constexpr (EXT) return iv.geo ? iv.reach : (int)floor(radius);
  else return (int)floor(radius);
}
template <bool EXT>
__device__ __forceinline__ void index_col_coords(const IndexView& iv, long long col, int& gx, int& gy) {
  if constexpr (EXT) { if (iv.geo) { gx = iv.cqx[col]; gy = iv.cqy[col]; } }
}
// level of a state level inside its variable (vertical localisation distances are per variable)
template <bool EXT>
__device__ __forceinline__ int index_level(const IndexView& iv, int lt) {
  if constexpr (EXT) return iv.levmap ? iv.levmap[lt] : lt;
  else return lt;
}

// Candidate range of one cell row for a column at global (gx, gy) and integer reach R = floor(r).
__device__ __forceinline__ void index_row_range(const IndexView& iv, int gx, int R, int cy,
                                                int& b, int& e) {
  int x0 = gx - R - iv.xmin, x1 = gx + R - iv.xmin;
  if (x1 < 0 || x0 >= iv.ncx * iv.cell) { b = e = 0; return; }
  int cx0 = max(x0, 0) / iv.cell, cx1 = min(x1 / iv.cell, iv.ncx - 1);
  b = iv.cell_start[cy * iv.ncx + cx0];
  e = iv.cell_start[cy * iv.ncx + cx1 + 1];
}
__device__ __forceinline__ void index_cy_range(const IndexView& iv, int gy, int R, int& cy0, int& cy1) {
  int y0 = gy - R - iv.ymin, y1 = gy + R - iv.ymin;
  if (y1 < 0 || y0 >= iv.ncy * iv.cell) { cy0 = 0; cy1 = -1; return; }
  cy0 = max(y0, 0) / iv.cell;
  cy1 = min(y1 / iv.cell, iv.ncy - 1);
}

// Location::haversine (Location.hpp:349-357) in the reference's operation order: explicit _rn arithmetic (no FMA
// contraction), CUDA's double sin / cos / atan2 (<= 2 ulp; glibc's are <= 1 ulp, so a distance can differ from the
// host's in the last bits -- the selection differs only for a pair within that of the radius).
__device__ __forceinline__ double geo_deg2rad(double deg) { return __ddiv_rn(__dmul_rn(deg, 3.14159265358979323846), 180.0); }
__device__ __noinline__ double geo_haversine(double lat1, double lon1, double lat2, double lon2) {
  const double dlat = geo_deg2rad(__dsub_rn(lat2, lat1)), dlon = geo_deg2rad(__dsub_rn(lon2, lon1));
  const double s1 = sin(__ddiv_rn(dlat, 2.0)), s2 = sin(__ddiv_rn(dlon, 2.0));
  const double a = __dadd_rn(__dmul_rn(s1, s1),
                             __dmul_rn(__dmul_rn(__dmul_rn(cos(geo_deg2rad(lat1)), cos(geo_deg2rad(lat2))), s2), s2));
  const double c = __dmul_rn(2.0, atan2(__dsqrt_rn(a), __dsqrt_rn(__dsub_rn(1.0, a))));
  return __dmul_rn(6371.0, c);
}
// out of line and fed by plain pointers (only EXT instantiations reference it)
__device__ __noinline__ double geo_distance(const double* __restrict__ clat, const double* __restrict__ clon,
                                            const double* __restrict__ slat, const double* __restrict__ slon,
                                            long long col, int a) {
  return geo_haversine(clat[col], clon[col], slat[a], slon[a]);
}

// Location::distance_to(...) <= radius (Location.hpp:204-230, LETKF.hpp:161-162).  GRID: bit-exact --
// dx, dy are exact integers, dx*dx+dy*dy is exact in FP64, sqrt is IEEE correctly rounded.
// GEOGRAPHIC: haversine kilometres between the column `col` and the observation at sorted position `a`.
template <bool EXT>
__device__ __forceinline__ bool index_within(const IndexView& iv, long long col, int a, int gx, int gy, double radius,
                                             double* dist_out) {
  double dist;
  bool geo = false;
  if constexpr (EXT) geo = iv.geo != 0;
  if (geo) {
    dist = geo_distance(iv.clat, iv.clon, iv.slat, iv.slon, col, a);
  } else {
    double dx = (double)(gx - iv.sx[a]), dy = (double)(gy - iv.sy[a]);
    dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  }
  *dist_out = dist;
  return dist <= radius;
}

// counts per owned column (tests: bit-exact against the brute-force oracle)
template <bool EXT>
__global__ void index_query_counts_kernel(IndexView iv, int nx, int own_nx, int own_ny, int gx0,
                                          int gy0, double radius, int32_t* __restrict__ counts) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)own_nx * own_ny) return;
  int lx = (int)(t % own_nx), ly = (int)(t / own_nx);
  int gx = gx0 + lx, gy = gy0 + ly;
  const long long col = (long long)ly * nx + lx;
  index_col_coords<EXT>(iv, col, gx, gy);
  int R = index_reach<EXT>(iv, radius);
  int cnt = 0;
  if (radius >= 0.0) {
    int cy0, cy1;
    index_cy_range(iv, gy, R, cy0, cy1);
    for (int cy = cy0; cy <= cy1; ++cy) {
      int b, e;
      index_row_range(iv, gx, R, cy, b, e);
      for (int a = b; a < e; ++a) {
        double dist;
        cnt += index_within<EXT>(iv, col, a, gx, gy, radius, &dist) ? 1 : 0;
      }
    }
  }
  counts[col] = cnt;
}

// explicit lists (global ids, kernel order) for selected columns
template <bool EXT>
__global__ void index_query_lists_kernel(IndexView iv, const int64_t* __restrict__ gid, int nx,
                                         int gx0, int gy0, double radius,
                                         const int64_t* __restrict__ cols, int64_t ncols, int cap,
                                         int64_t* __restrict__ lists, int32_t* __restrict__ counts) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncols) return;
  int64_t col = cols[t];
  int gx = gx0 + (int)(col % nx), gy = gy0 + (int)(col / nx);
  index_col_coords<EXT>(iv, col, gx, gy);
  int R = index_reach<EXT>(iv, radius);
  int cnt = 0;
  if (radius >= 0.0) {
    int cy0, cy1;
    index_cy_range(iv, gy, R, cy0, cy1);
    for (int cy = cy0; cy <= cy1; ++cy) {
      int b, e;
      index_row_range(iv, gx, R, cy, b, e);
      for (int a = b; a < e; ++a) {
        double dist;
        if (index_within<EXT>(iv, col, a, gx, gy, radius, &dist)) {
          if (cnt < cap) lists[t * cap + cnt] = gid[iv.sorted_row[a]];
          ++cnt;
        }
      }
    }
  }
  counts[t] = cnt;
}
