// H(x): 4-point inverse-distance interpolation of every member at every observation, fused with
// the observation-space mean / perturbation / innovation epilogue.
//
// Replaces k calls of IdentityObsOperator::apply (IdentityObsOperator.hpp:154-180) with
// find4NearestGridPoints (:594-638) + idw4Interpolation (:643-676), and LETKF.hpp:209-211.
// One warp per observation, lanes over members (the [col][lev][member] layout makes the four
// corner reads 4 contiguous runs of k doubles).  All FP64 arithmetic uses explicit _rn
// intrinsics in the reference's operation order, so Y, ybar, Y' and d are bit-identical to the
// reference's plain (non-FMA) build.  HBM/L2-bound: 4*k*8 B read + 2*k*8 B written per obs.
#pragma once
#include "mdc_internal.cuh"

#define MDC_MAX_VARS 16
struct HxGeom {
  int nx, ny, nz, k;          // local grid; nz = all levels of all variables
  int gx0, gy0, gnx, gny;     // placement in global grid
  // multi-variable states (WRF-shaped, [var][lev][y][x] per member): the neighbour search runs on the geometry's
  // nzg levels (IdentityObsOperator.hpp:684-690), the value is read from the observation's variable, whose 2-D
  // fields ignore the level (:701-711).  nvar = 0: one variable of nz levels.
  int nzg, nvar;
  int var_off[MDC_MAX_VARS], var_nlev[MDC_MAX_VARS];
};

__device__ __forceinline__ double hx_dist2(double x, double y, int ii, int jj) {
  double dx = __dsub_rn(x, (double)ii), dy = __dsub_rn(y, (double)jj);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
__device__ __forceinline__ double hx_dist3(double x, double y, double z, int ii, int jj, int kk) {
  double dx = __dsub_rn(x, (double)ii), dy = __dsub_rn(y, (double)jj), dz = __dsub_rn(z, (double)kk);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

template <int WARPS>
__global__ void hx_idw4_kernel(const double* __restrict__ X, HxGeom g, int64_t P,
                               const int32_t* __restrict__ ox, const int32_t* __restrict__ oy,
                               const int32_t* __restrict__ oz, const int32_t* __restrict__ ovar,
                               const uint8_t* __restrict__ valid,
                               const double* __restrict__ oval, double* __restrict__ Y,
                               double* __restrict__ ybar, double* __restrict__ Yp,
                               double* __restrict__ d, int* __restrict__ err_flag) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* row = sm + (size_t)warp * g.k;
  for (int64_t i = (int64_t)blockIdx.x * WARPS + warp; i < P; i += (int64_t)gridDim.x * WARPS) {
    const bool ok = valid[i] != 0;
    int ii[8], jj[8], kk[8];
    double w[4];
    int64_t base[4];
    double wsum = 0.0;
    if (ok) {
      // clamp to the GLOBAL grid bounds (:598-600)
      double x = fmax(0.0, fmin((double)(g.gnx - 1), (double)ox[i]));
      double y = fmax(0.0, fmin((double)(g.gny - 1), (double)oy[i]));
      double z = fmax(0.0, fmin((double)(g.nzg - 1), (double)oz[i]));
      int voff = 0, vn = g.nz;
      if (g.nvar > 0) { const int v = ovar ? ovar[i] : 0; voff = g.var_off[v]; vn = g.var_nlev[v]; }
      int i0 = (int)floor(x), j0 = (int)floor(y), k0 = (int)floor(z);
      int i1 = min(i0 + 1, g.gnx - 1), j1 = min(j0 + 1, g.gny - 1), k1 = min(k0 + 1, g.nzg - 1);
      double dist[8];
      if (g.nzg == 1) {
        ii[0] = i0; jj[0] = j0; ii[1] = i1; jj[1] = j0; ii[2] = i0; jj[2] = j1; ii[3] = i1; jj[3] = j1;
#pragma unroll
        for (int c = 0; c < 4; ++c) { kk[c] = 0; dist[c] = hx_dist2(x, y, ii[c], jj[c]); }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          ii[c] = (c & 1) ? i1 : i0; jj[c] = (c & 2) ? j1 : j0; kk[c] = (c & 4) ? k1 : k0;
          dist[c] = hx_dist3(x, y, z, ii[c], jj[c], kk[c]);
        }
        // stable insertion sort by distance (std::sort on 8 elements, :621-624), keep 4
#pragma unroll
        for (int a = 1; a < 8; ++a) {
#pragma unroll
          for (int b = a; b > 0; --b) {
            if (dist[b] < dist[b - 1]) {
              double td = dist[b]; dist[b] = dist[b - 1]; dist[b - 1] = td;
              int t = ii[b]; ii[b] = ii[b - 1]; ii[b - 1] = t;
              t = jj[b]; jj[b] = jj[b - 1]; jj[b - 1] = t;
              t = kk[b]; kk[b] = kk[b - 1]; kk[b - 1] = t;
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        w[c] = (dist[c] == 0.0) ? 1e12 : __ddiv_rn(1.0, dist[c]);
        wsum = __dadd_rn(wsum, w[c]);
        int lx = ii[c] - g.gx0, ly = jj[c] - g.gy0;
        if (lx < 0 || lx >= g.nx || ly < 0 || ly >= g.ny) {
          if (lane == 0) atomicExch(err_flag, 1);   // obs needs state outside this rank's tile+halo
          lx = max(0, min(g.nx - 1, lx)); ly = max(0, min(g.ny - 1, ly));
        }
        base[c] = (((int64_t)ly * g.nx + lx) * g.nz + voff + (vn > 1 ? kk[c] : 0)) * g.k;
      }
    }
    for (int m = lane; m < g.k; m += 32) {
      double h = 0.0;
      if (ok) {
        double ws = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) ws = __dadd_rn(ws, __dmul_rn(w[c], X[base[c] + m]));
        h = __ddiv_rn(ws, wsum);
      }
      row[m] = h;
    }
    __syncwarp();
    double s = 0.0;
    if (lane == 0) {
      for (int m = 0; m < g.k; ++m) s = __dadd_rn(s, row[m]);   // rowwise().mean(): sum in order / k
      s = __ddiv_rn(s, (double)g.k);
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    for (int m = lane; m < g.k; m += 32) {
      double h = row[m];
      if (Y) Y[i * g.k + m] = h;
      Yp[i * g.k + m] = __dsub_rn(h, s);
    }
    if (lane == 0) {
      ybar[i] = s;
      d[i] = __dsub_rn(oval[i], s);
    }
    __syncwarp();
  }
}
