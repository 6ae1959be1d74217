// GEOGRAPHIC observations and multi-variable states (the WRF-shaped case).
//
//   geo_locate_kernel   replaces IdentityObsOperator::convertGeographicToGrid (IdentityObsOperator.hpp:484-530):
//                       nearest grid point of every observation = FIRST minimum over the grid's 2-D latitude /
//                       longitude arrays of sqrt((lon - glon)^2 + (lat - glat)^2) (degrees, strict '<'), nearest
//                       vertical level = first minimum of |level - vertical_coords[z]|.  Bit-exact: same
//                       operation order, _rn arithmetic, IEEE sqrt.  The grid is streamed through shared memory
//                       in tiles (every thread of a block reads the same point: broadcast); the square root is
//                       only taken when the squared distance improves (sqrt is monotone, so the strict-'<'
//                       decision on the rounded roots is unchanged).
//   geo_quantise_kernel lattice coordinates for the bucket index of a haversine selection
//                       (Location.hpp:213-217, 349-357): longitudes are unwrapped about the domain centre, both
//                       axes are divided into quanta of 1/GEO_SUB of the largest latitude / longitude difference
//                       an observation inside the radius can have, so the integer reach GEO_SUB + 2 is
//                       conservative and the cell walk of index_kernels.cuh applies unchanged.  Observations far
//                       outside the domain are parked on a strip of cells beyond every column's reach.
//   geo_gather_sorted_kernel  true coordinates in index order, read by geo_distance() in the column kernels.
//
// Integer / byte work next to a handful of FP64 operations: HBM- and shared-memory-bound, off the column kernel's path.
#pragma once
#include <float.h>

#include "mdc_internal.cuh"

#define GEO_SUB 8          // lattice quanta per selection reach
#define GEO_FAR 256        // cells of the strip that parks the observations beyond every column's reach
#define GEO_TILE 2048      // grid points staged per shared-memory tile (32 KB)

__global__ void __launch_bounds__(256) geo_locate_kernel(int64_t P, const double* __restrict__ olat,
                                                         const double* __restrict__ olon,
                                                         const double* __restrict__ olev,
                                                         const double* __restrict__ glat,
                                                         const double* __restrict__ glon, int64_t G, int nx,
                                                         const double* __restrict__ vcoord, int nlev,
                                                         int32_t* __restrict__ ox, int32_t* __restrict__ oy,
                                                         int32_t* __restrict__ oz) {
  __shared__ double s_lat[GEO_TILE], s_lon[GEO_TILE];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < P;
  const double lat = act ? olat[i] : 0.0, lon = act ? olon[i] : 0.0;
  double best2 = INFINITY, best = DBL_MAX;   // min_dist = numeric_limits<double>::max() (:494)
  int64_t best_idx = 0;
  for (int64_t base = 0; base < G; base += GEO_TILE) {
    const int n = (int)min((int64_t)GEO_TILE, G - base);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) { s_lat[t] = glat[base + t]; s_lon[t] = glon[base + t]; }
    __syncthreads();
    if (act) {
#pragma unroll 4
      for (int t = 0; t < n; ++t) {
        const double dx = __dsub_rn(lon, s_lon[t]), dy = __dsub_rn(lat, s_lat[t]);
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));   // (:502-503)
        if (d2 < best2) {
          const double dd = __dsqrt_rn(d2);
          if (dd < best) { best = dd; best2 = d2; best_idx = base + t; }    // strict '<': first minimum (:505-508)
        }
      }
    }
  }
  if (!act) return;
  int kz = 0;
  if (vcoord && nlev > 0) {                                                  // (:515-526)
    const double level = olev ? olev[i] : 0.0;
    double mv = DBL_MAX;
    for (int z = 0; z < nlev; ++z) {
      const double dist = fabs(__dsub_rn(level, vcoord[z]));
      if (dist < mv) { mv = dist; kz = z; }
    }
  }
  ox[i] = (int32_t)(best_idx % nx);
  oy[i] = (int32_t)(best_idx / nx);
  oz[i] = kz;
}

// ---- O(P) nearest grid point: the grid points are bucketed once per geometry into square cells of the raw
// (longitude, latitude) plane (the reference's metric is the plain Euclidean distance in degrees, no wrap) at two
// resolutions -- ~4 points per fine cell, 16 x 16 fine cells per coarse cell.  An observation walks the rings of
// cells around its own (clamped into the cell grid when it lies outside).  Before ring rho every unvisited point
// lies in a cell at Chebyshev distance >= rho, so it is at least
//     LB(rho) = sqrt(min((dx_out + (rho-1) c)^2 + dy_out^2, dx_out^2 + (dy_out + (rho-1) c)^2))
// away, dx_out / dy_out being the observation's distance to the grid's bounding box along each axis (0 inside); the
// walk stops once the best distance is below LB(rho) (1 - 1e-9) -- the margin covers the rounding of the cell
// assignment.  Observations that are not settled within GEO_FINE_RINGS fine rings (far outside the domain, or in an
// empty corner of a projected grid's bounding box) finish on the coarse cells, whose rings grow 16 times faster.
// Ties are resolved as the reference's linear scan does (strict '<' in index order = smallest linear index among
// equal rounded distances), so the result is bit-identical to geo_locate_kernel.
#define GEO_FINE_RINGS 4
#define GEO_COARSE 16
struct GeoCells {
  double lon0, lat0, lon1, lat1;   // bounding box of the grid points
  double inv_c, c;
  int ncx, ncy;
  const int32_t *start, *pts;      // [ncx*ncy + 1], [G] linear grid index in cell order
  const double *plat, *plon;       // [G] coordinates in cell order
};

__device__ __forceinline__ void geo_cell_of(const GeoCells& gc, double lat, double lon, int& cx, int& cy) {
  const double fx = floor((lon - gc.lon0) * gc.inv_c), fy = floor((lat - gc.lat0) * gc.inv_c);
  cx = (int)fmin(fmax(fx, 0.0), (double)(gc.ncx - 1));
  cy = (int)fmin(fmax(fy, 0.0), (double)(gc.ncy - 1));
}

__global__ void geo_cell_key_kernel(int64_t G, const double* __restrict__ glat, const double* __restrict__ glon,
                                    GeoCells gc, int32_t* __restrict__ key, int32_t* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (int64_t)gridDim.x * blockDim.x) {
    int cx, cy;
    geo_cell_of(gc, glat[i], glon[i], cx, cy);
    const int kk = cy * gc.ncx + cx;
    key[i] = kk;
    atomicAdd(hist + kk, 1);
  }
}

// coordinates in cell order next to the linear grid index: the ring walk then reads three sequential arrays
__global__ void geo_cell_gather_kernel(int64_t G, const int32_t* __restrict__ pts, const double* __restrict__ glat,
                                       const double* __restrict__ glon, double* __restrict__ plat,
                                       double* __restrict__ plon) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < G; a += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = pts[a];
    plat[a] = glat[g]; plon[a] = glon[g];
  }
}

// rings 0 .. rho_cap of one cell level; true when the search is settled (bound met or every cell visited)
__device__ __forceinline__ bool geo_ring_walk(const GeoCells& gc, double lat, double lon, int rho_cap, double& best,
                                              double& best2, int& best_idx) {
  int cx, cy;
  geo_cell_of(gc, lat, lon, cx, cy);
  const double dx_out = fmax(0.0, fmax(gc.lon0 - lon, lon - gc.lon1));
  const double dy_out = fmax(0.0, fmax(gc.lat0 - lat, lat - gc.lat1));
  const int rho_max = max(max(cx, gc.ncx - 1 - cx), max(cy, gc.ncy - 1 - cy));
  for (int rho = 0; rho <= rho_max; ++rho) {
    if (rho > 1) {
      const double w = (double)(rho - 1) * gc.c;
      const double lb2 = fmin((dx_out + w) * (dx_out + w) + dy_out * dy_out, dx_out * dx_out + (dy_out + w) * (dy_out + w));
      if (best < sqrt(lb2) * (1.0 - 1e-9)) return true;
    }
    if (rho > rho_cap) return false;
    const int y0 = cy - rho, y1 = cy + rho;
    for (int yy = max(y0, 0); yy <= min(y1, gc.ncy - 1); ++yy) {
      // the whole row on the ring's top / bottom edge, its two end cells otherwise
      const bool edge_row = (yy == y0 || yy == y1);
      const int step = edge_row ? 1 : max(2 * rho, 1);
      int xx = cx - rho;
      if (edge_row && xx < 0) xx = 0;
      for (; xx <= min(cx + rho, gc.ncx - 1); xx += step) {
        if (xx < 0) continue;
        const int cell = yy * gc.ncx + xx;
        const int b = gc.start[cell], e = gc.start[cell + 1];
        for (int a = b; a < e; ++a) {
          const double dx = __dsub_rn(lon, gc.plon[a]), dy = __dsub_rn(lat, gc.plat[a]);
          const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (d2 <= best2 * (1.0 + 1e-15)) {   // (a larger d2 can still round to the same root and win on its index)
            const double dd = __dsqrt_rn(d2);
            const int g = gc.pts[a];
            if (dd < best || (dd == best && g < best_idx)) { best = dd; best2 = d2; best_idx = g; }
          }
        }
      }
    }
  }
  return true;
}

__global__ void __launch_bounds__(128) geo_locate_ring_kernel(int64_t P, const double* __restrict__ olat,
                                                              const double* __restrict__ olon,
                                                              const double* __restrict__ olev, GeoCells fine,
                                                              GeoCells coarse, int nx,
                                                              const double* __restrict__ vcoord, int nlev,
                                                              int32_t* __restrict__ ox, int32_t* __restrict__ oy,
                                                              int32_t* __restrict__ oz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double lat = olat[i], lon = olon[i];
  double best = DBL_MAX, best2 = INFINITY;
  int best_idx = 0x7fffffff;
  if (!geo_ring_walk(fine, lat, lon, GEO_FINE_RINGS, best, best2, best_idx))
    geo_ring_walk(coarse, lat, lon, 0x7fffffff, best, best2, best_idx);
  int kz = 0;
  if (vcoord && nlev > 0) {
    const double level = olev ? olev[i] : 0.0;
    double mv = DBL_MAX;
    for (int z = 0; z < nlev; ++z) {
      const double dist = fabs(__dsub_rn(level, vcoord[z]));
      if (dist < mv) { mv = dist; kz = z; }
    }
  }
  ox[i] = best_idx % nx;
  oy[i] = best_idx / nx;
  oz[i] = kz;
}

struct GeoLattice {
  double lon_c;              // unwrap centre: u = (lon - lon_c) - 360 rint((lon - lon_c) / 360) in [-180, 180]
  double u0, lat0;           // lattice origin (minimum over the columns)
  double inv_qx, inv_qy;     // 1 / quantum
  double lo_x, hi_x, lo_y, hi_y;   // lattice bounds (columns +- more than the reach); outside: the far strip
};

__device__ __forceinline__ void geo_lattice_coords(const GeoLattice& g, double lat, double lon, int64_t i, int& qx, int& qy) {
  const double u = (lon - g.lon_c) - 360.0 * rint((lon - g.lon_c) / 360.0);
  const double fx = floor((u - g.u0) * g.inv_qx), fy = floor((lat - g.lat0) * g.inv_qy);
  if (fx < g.lo_x || fx > g.hi_x || fy < g.lo_y || fy > g.hi_y) {
    // beyond every column's reach: parked on a strip of GEO_FAR cells left of the lattice (spread by the row
    // number so that no single cell -- one thread of the per-cell sort -- collects them all)
    qx = (int)g.lo_x - (GEO_SUB + 2) * (1 + (int)(i & (GEO_FAR - 1)));
    qy = (int)g.lo_y;
    return;
  }
  qx = (int)fx; qy = (int)fy;
}

__global__ void geo_quantise_kernel(int64_t n, const double* __restrict__ lat, const double* __restrict__ lon,
                                    GeoLattice g, int32_t* __restrict__ qx, int32_t* __restrict__ qy) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int a, b;
    geo_lattice_coords(g, lat[i], lon[i], i, a, b);
    qx[i] = a; qy[i] = b;
  }
}

__global__ void geo_gather_sorted_kernel(int64_t P, const int32_t* __restrict__ sorted_row,
                                         const double* __restrict__ lat, const double* __restrict__ lon,
                                         double* __restrict__ slat, double* __restrict__ slon) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < P; a += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = sorted_row[a];
    slat[a] = lat[r]; slon[a] = lon[r];
  }
}
