// lwenkf_kernels.cuh -- device kernels of the locally weighted EnKF (LWEnKF.hpp:207-334).
//
// The reference forms everything densely on the host: the perturbation matrix, S = (sum_m w_m y'_m y'_m^T) o L + R
// (P x P), the gain K = X' Y'^T S^-1 / (k - 1) (n x P), its Schur product with a second localisation matrix, and
// K (yo + eps_m - Yb_m) per member.  Both "localisations" use INDEX distances |i - j| / dim (the reference's own
// simplification, LWEnKF.hpp:566-570, 589-594), which is why the gain cannot be kept in ensemble space: the Schur
// product with Lg is applied entry by entry.  Here S is built and factorised on the device (LU with partial
// pivoting: S need not be positive definite for the cutoff / polynomial localisation functions), K is never stored:
// a tiled kernel forms K o Lg one (state points x observations) tile at a time in shared memory, applies it to the
// innovation matrix and streams max / min of K.
#pragma once

// LWEnKF.hpp:603-636 (MDC_LOC_* codes: 0 cutoff, 2 gaussian, 3 exponential, 4 the reference's polynomial)
__device__ __forceinline__ double lw_loc_fn(int fn, double distance, double radius) {
  const double nd = distance / radius;
  switch (fn) {
    case MDC_LOC_GAUSSIAN: return exp(-0.5 * nd * nd);
    case MDC_LOC_EXPONENTIAL: return exp(-nd);
    case MDC_LOC_CUTOFF: return nd <= 1.0 ? 1.0 : 0.0;
    case MDC_LOC_REF_GASPARI_COHN:
      if (nd >= 2.0) return 0.0;
      if (nd >= 1.0) { const double z = nd - 1.0; return ((-0.25 * z + 0.5) * z + 0.625) * z + 0.125; }
      return (((-0.25 * nd + 0.5) * nd + 0.625) * nd - 5.0) * nd + 4.0;
    default: return exp(-0.5 * nd * nd);
  }
}

// per-member sum of squared perturbations: partial[block][m] (fixed order -> deterministic); X is [point][member]
__global__ void lw_member_sqnorm_kernel(const double* __restrict__ X, const double* __restrict__ mean, int64_t npts, int k,
                                        double* __restrict__ partial) {
  __shared__ double sh[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t p = (int64_t)blockIdx.x * nw + warp; p < npts; p += (int64_t)gridDim.x * nw) {
    const double mu = mean[p];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = lane + 32 * q;
      if (m < k) { const double v = X[p * k + m] - mu; acc[q] = fma(v, v, acc[q]); }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { const int m = lane + 32 * q; if (m < k) sh[warp][m] = acc[q]; }
  __syncthreads();
  for (int m = threadIdx.x; m < k; m += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += sh[w][m];
    partial[(int64_t)blockIdx.x * k + m] = s;
  }
}

// likelihood exponent per member: q_m = sum_a (yo_a - Y_am)^2 / var_a (invalid observations: weight 0)
__global__ void lw_likelihood_kernel(const double* __restrict__ Y, const double* __restrict__ val, const double* __restrict__ err,
                                     const uint8_t* __restrict__ valid, int64_t P, int k, double* __restrict__ q) {
  __shared__ double sh[32];
  const int m = blockIdx.x;
  double s = 0.0;
  for (int64_t a = threadIdx.x; a < P; a += blockDim.x) {
    const double in = val[a] - Y[a * k + m];
    if (valid[a]) s += in * (in / (err[a] * err[a]));
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    q[m] = t;
  }
}

// S (column-major = row-major, symmetric in exact arithmetic) = (sum_m w_m y'_am y'_bm) L(a, b) + delta_ab var_a
__global__ void lw_build_S_kernel(const double* __restrict__ Yp, const double* __restrict__ w, const double* __restrict__ err,
                                  const uint8_t* __restrict__ valid, int64_t P, int k, int loc_fn, double radius,
                                  double* __restrict__ S) {
  extern __shared__ double shw[];
  for (int m = threadIdx.x; m < k; m += blockDim.x) shw[m] = w[m];
  __syncthreads();
  const int64_t total = P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = e / P, b = e - a * P;
    double s = 0.0;
    for (int m = 0; m < k; ++m) s += shw[m] * Yp[a * k + m] * Yp[b * k + m];
    s *= lw_loc_fn(loc_fn, fabs((double)(a - b)) / (double)P, radius);
    if (a == b) s += valid[a] ? err[a] * err[a] : INFINITY;
    S[e] = s;
  }
}

// [P][k] row-major <-> column-major with leading dimension P (the LAPACK-style solver's right-hand sides)
__global__ void lw_transpose_kernel(const double* __restrict__ in, int64_t rows, int cols, double* __restrict__ out, int to_colmajor) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols;
    const int c = (int)(e - r * cols);
    if (to_colmajor) out[(int64_t)c * rows + r] = in[e];
    else out[e] = in[(int64_t)c * rows + r];
  }
}

// The update, K never stored.  Gs = S^-1 Y' [P][k], D = yo + eps - Yb [P][k].  Per block: LW_TP state points; per
// observation tile of LW_TO: Kt[p][j] = Lg(i_p, j) (x'_p . Gs_j) / (k - 1), out[p][m] += Kt[p][j] D[j][m].
// The state index of the reference's vectors is the host order [lev][y][x]: i = lev G + col for the device point
// col nz + lev.  partial_mm[block][2] = max / min of K over the block's rows.
#define LW_TP 32
#define LW_TO 32
__global__ void __launch_bounds__(256) lw_apply_kernel(double* __restrict__ X, const double* __restrict__ mean,
                                                       const double* __restrict__ Gs, const double* __restrict__ D, int64_t npts,
                                                       int64_t P, int k, int nz, int64_t G, double sqrt_infl, int loc_fn,
                                                       double radius, double* __restrict__ partial_mm) {
  extern __shared__ double sm[];
  const int ks = k | 1;
  double* xp = sm;                       // [LW_TP][ks] inflated perturbations
  double* gs = xp + LW_TP * ks;          // [LW_TO][ks]
  double* dd = gs + LW_TO * ks;          // [LW_TO][ks]
  double* kt = dd + LW_TO * ks;          // [LW_TP][LW_TO + 1]
  double* mu = kt + LW_TP * (LW_TO + 1); // [LW_TP]
  __shared__ double rmax[8], rmin[8];
  const int tid = threadIdx.x;
  const int64_t dim = npts > P ? npts : P;
  const double inv_km1 = 1.0 / (double)(k - 1);
  double kmax = -INFINITY, kmin = INFINITY;
  for (int64_t p0 = (int64_t)blockIdx.x * LW_TP; p0 < npts; p0 += (int64_t)gridDim.x * LW_TP) {
    const int np = (int)min((int64_t)LW_TP, npts - p0);
    if (tid < np) mu[tid] = mean[p0 + tid];
    __syncthreads();
    for (int e = tid; e < np * k; e += blockDim.x) {
      const int p = e / k, m = e - p * k;
      xp[p * ks + m] = (X[(p0 + p) * k + m] - mu[p]) * sqrt_infl;
    }
    // accumulators: thread handles entries e = tid, tid + 256, ... of the [np][k] output tile
    double acc[(LW_TP * 128 + 255) / 256];
#pragma unroll
    for (int q = 0; q < (LW_TP * 128 + 255) / 256; ++q) acc[q] = 0.0;
    for (int64_t j0 = 0; j0 < P; j0 += LW_TO) {
      const int no = (int)min((int64_t)LW_TO, P - j0);
      __syncthreads();
      for (int e = tid; e < no * k; e += blockDim.x) {
        const int j = e / k, m = e - j * k;
        gs[j * ks + m] = Gs[(j0 + j) * k + m];
        dd[j * ks + m] = D[(j0 + j) * k + m];
      }
      __syncthreads();
      for (int e = tid; e < np * no; e += blockDim.x) {
        const int p = e / no, j = e - p * no;
        double s = 0.0;
        for (int m = 0; m < k; ++m) s = fma(xp[p * ks + m], gs[j * ks + m], s);
        const int64_t pt = p0 + p, col = pt / nz;
        const int64_t i = (pt - col * nz) * G + col;
        s = s * inv_km1 * lw_loc_fn(loc_fn, fabs((double)(i - (j0 + j))) / (double)dim, radius);
        kt[p * (LW_TO + 1) + j] = s;
        kmax = fmax(kmax, s);
        kmin = fmin(kmin, s);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < (LW_TP * 128 + 255) / 256; ++q) {
        const int e = tid + q * 256;
        if (e < np * k) {
          const int p = e / k, m = e - p * k;
          double s = acc[q];
          for (int j = 0; j < no; ++j) s = fma(kt[p * (LW_TO + 1) + j], dd[j * ks + m], s);
          acc[q] = s;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < (LW_TP * 128 + 255) / 256; ++q) {
      const int e = tid + q * 256;
      if (e < np * k) {
        const int p = e / k, m = e - p * k;
        X[(p0 + p) * k + m] = mu[p] + (xp[p * ks + m] + acc[q]);
      }
    }
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) {
    kmax = fmax(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    kmin = fmin(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
  }
  if ((tid & 31) == 0) { rmax[tid >> 5] = kmax; rmin[tid >> 5] = kmin; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { kmax = fmax(kmax, rmax[w]); kmin = fmin(kmin, rmin[w]); }
    partial_mm[2 * blockIdx.x] = kmax;
    partial_mm[2 * blockIdx.x + 1] = kmin;
  }
}
