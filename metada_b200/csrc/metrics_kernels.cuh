// metrics_kernels.cuh -- ensemble verification metrics against a truth state on the device
// (reference: framework/algorithms/Metrics.hpp:74-290 -- mean, spread, RMSE, bias, correlation,
// CRPS, average spread).  One warp per state point: the k member values of a point are contiguous
// in the [col][lev][member] layout, so a warp reads them coalesced and keeps them in registers.
// Partial sums are combined in a fixed order (per-warp registers -> per-block slot -> one final
// block), so the result does not depend on scheduling.
#pragma once
#include "mdc_internal.cuh"

#define MT_NSUM 9   // sum (m-t), (m-t)^2, m, t, m t, m^2, t^2, spread, crps point terms
#define MT_MAXR 4   // members per lane (k <= 128)

__global__ void __launch_bounds__(256) metrics_points_kernel(const double* __restrict__ X,
                                                             const double* __restrict__ truth, int64_t npoints,
                                                             int k, double* __restrict__ spread_out,
                                                             double* __restrict__ partial /*[grid][MT_NSUM]*/) {
  __shared__ double sh[8][MT_NSUM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nr = (k + 31) >> 5;
  double acc[MT_NSUM];
#pragma unroll
  for (int q = 0; q < MT_NSUM; ++q) acc[q] = 0.0;
  const double rk = 1.0 / (double)k;
  for (int64_t pt = (int64_t)blockIdx.x * 8 + warp; pt < npoints; pt += (int64_t)gridDim.x * 8) {
    const double* x = X + pt * k;
    double v[MT_MAXR];
#pragma unroll
    for (int r = 0; r < MT_MAXR; ++r) v[r] = (r < nr && lane + 32 * r < k) ? x[lane + 32 * r] : 0.0;
    const double t = truth[pt];
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < MT_MAXR; ++r) s += v[r];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double m = s * rk;                                     // Metrics.hpp:108-121
    double d2 = 0.0, at = 0.0, pair = 0.0;
#pragma unroll
    for (int r = 0; r < MT_MAXR; ++r)
      if (r < nr && lane + 32 * r < k) {
        const double d = v[r] - m;
        d2 = fma(d, d, d2);                                      // :137-150
        at += fabs(v[r] - t);                                    // :246-249
      }
    // sum_j sum_l |y_j - y_l| (:240-245): every member is broadcast in turn
    for (int r2 = 0; r2 < nr; ++r2) {
      const int cnt = min(32, k - 32 * r2);
      for (int l = 0; l < cnt; ++l) {
        const double y = __shfl_sync(0xffffffffu, v[r2 < MT_MAXR ? r2 : 0], l);
#pragma unroll
        for (int r = 0; r < MT_MAXR; ++r)
          if (r < nr && lane + 32 * r < k) pair += fabs(v[r] - y);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      at += __shfl_xor_sync(0xffffffffu, at, o);
      pair += __shfl_xor_sync(0xffffffffu, pair, o);
    }
    if (lane == 0) {
      const double sp = sqrt(d2 / (double)(k - 1));
      if (spread_out) spread_out[pt] = sp;
      const double e = m - t;
      acc[0] += e; acc[1] = fma(e, e, acc[1]); acc[2] += m; acc[3] += t;
      acc[4] = fma(m, t, acc[4]); acc[5] = fma(m, m, acc[5]); acc[6] = fma(t, t, acc[6]);
      acc[7] += sp;
      acc[8] += at * rk - pair / (2.0 * (double)k * (double)k);  // :250-251
    }
  }
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < MT_NSUM; ++q) sh[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < MT_NSUM) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    partial[(size_t)blockIdx.x * MT_NSUM + threadIdx.x] = s;
  }
}

__global__ void metrics_final_kernel(const double* __restrict__ partial, int nblocks, double npoints,
                                     double* __restrict__ out /*[5]: rmse bias correlation crps avg_spread*/) {
  __shared__ double tot[MT_NSUM];
  if (threadIdx.x < MT_NSUM) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * MT_NSUM + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = npoints;
    out[0] = sqrt(tot[1] / n);                                                                 // :265-273
    out[1] = tot[0] / n;                                                                       // :166-173
    out[2] = (n * tot[4] - tot[2] * tot[3]) /
             sqrt((n * tot[5] - tot[2] * tot[2]) * (n * tot[6] - tot[3] * tot[3]));            // :189-210
    out[3] = tot[8] / n;                                                                       // :253
    out[4] = tot[7] / n;                                                                       // :286-292
  }
}
