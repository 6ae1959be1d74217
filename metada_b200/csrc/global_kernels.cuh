// Non-localised (global) analysis kernels: ETKF (ETKF.hpp:100-179) and the stochastic EnKF
// (EnKF.hpp:139-256).
//
// Both reduce to ensemble space.  With C = Y'^T R^-1 Y', A = (k-1) I + C  (k x k):
//   ETKF : Pa = A^-1, wa = Pa Y'^T R^-1 d, Wa = sqrt(k-1) chol(Pa), Xa = xbar + infl X' (wa 1^T + Wa)
//   EnKF : S = Y'Y'^T/(k-1) + R is P x P, but by the Woodbury identity
//            Y'^T S^-1 = (k-1) A^-1 Y'^T R^-1
//          so K = X' Y'^T S^-1/(k-1) = X' A^-1 Y'^T R^-1 exactly, and
//            Xa = xbar 1^T + sqrt(infl) X' (I + A^-1 F),  F = Y'^T R^-1 D,  D = yo 1^T + eps - Y.
//          K (n x P) is never stored: its max/min are streamed out of a tiled FP64 GEMM.
// Kernels: obs-space Gram reductions (HBM-bound over Y'), a one-CTA k x k solve, a streaming
// state-update GEMM (n x k x k), and the gain min/max GEMM (n x P x k, FP64-pipe bound).
#pragma once
#include "letkf_kernels.cuh"
#include "mdc_internal.cuh"

#define GK_THREADS 256
#define GK_ROWS 32

// partial[b][0 .. k*k) = sum_rows rinv * a a^T ; partial[b][k*k .. k*k + k*kb) = sum_rows rinv * a b^T
__global__ void __launch_bounds__(GK_THREADS)
obs_gram_kernel(const double* __restrict__ Yp, const double* __restrict__ B, int kb,
                const double* __restrict__ err, const uint8_t* __restrict__ valid, int64_t P, int k,
                double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* Ar = sm;                         // [GK_ROWS][k]  (pre-multiplied by rinv)
  double* Ac = Ar + (size_t)GK_ROWS * k;   // [GK_ROWS][k]
  double* Br = Ac + (size_t)GK_ROWS * k;   // [GK_ROWS][kb]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nout = k * k + k * kb;
  double* out = partial + (size_t)blockIdx.x * nout;
  for (int e = tid; e < nout; e += nt) out[e] = 0.0;
  // contiguous slice of rows per block keeps the summation order fixed
  const int64_t per = (P + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(P, r0 + per);
  for (int64_t base = r0; base < r1; base += GK_ROWS) {
    const int rows = (int)min((int64_t)GK_ROWS, r1 - base);
    for (int e = tid; e < rows * k; e += nt) {
      const int r = e / k;
      const int64_t i = base + r;
      const double ee = err[i];
      const double rinv = valid[i] ? 1.0 / (ee * ee) : 0.0;
      const double v = Yp[i * k + (e - r * k)];
      Ac[e] = v;
      Ar[e] = v * rinv;
    }
    // (an invalid observation has weight 0; its innovation may be NaN -- a missing value -- and 0 * NaN would spread)
    for (int e = tid; e < rows * kb; e += nt) Br[e] = valid[base + e / kb] ? B[(base + e / kb) * kb + e % kb] : 0.0;
    __syncthreads();
    for (int e = tid; e < nout; e += nt) {
      double s = out[e];
      if (e < k * k) {
        const int a = e / k, b = e - a * k;
        for (int r = 0; r < rows; ++r) s += Ar[r * k + a] * Ac[r * k + b];
      } else {
        const int e2 = e - k * k, a = e2 / kb, b = e2 - a * kb;
        for (int r = 0; r < rows; ++r) s += Ar[r * k + a] * Br[r * kb + b];
      }
      out[e] = s;
    }
    __syncthreads();
  }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nb, int len,
                                       double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double s = 0.0;
  for (int b = 0; b < nb; ++b) s += partial[(size_t)b * len + e];
  out[e] = s;
}

// D[i][m] = (yo_i + sqrt(var_i) Z[i][m]) - Y[i][m]   (EnKF.hpp:219-222, 340-361)
__global__ void enkf_innov_kernel(const double* __restrict__ Y, const double* __restrict__ Z,
                                  const double* __restrict__ val, const double* __restrict__ err,
                                  int64_t P, int k, uint64_t seed, double* __restrict__ D) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P * k;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / k;
    double z;
    if (Z) z = Z[e];
    else {  // counter-based Box-Muller
      uint64_t h1 = mdc_hash(seed, 2 * (uint64_t)e), h2 = mdc_hash(seed, 2 * (uint64_t)e + 1);
      double u1 = ((double)(h1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);
      double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);
      z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    const double sd = sqrt(err[i] * err[i]);
    D[e] = (val[i] + sd * z) - Y[e];
  }
}

// One CTA: the k x k ensemble-space solve.  G = [C (k*k) | F or g].
//   mode 0 (ETKF): Wout = infl * (wa 1^T + sqrt(k-1) chol((C + (k-1)I)^-1))
//   mode 1 (EnKF): Wout = sqrt(infl) * (I + A^-1 F),  Ainv_out = A^-1
__global__ void __launch_bounds__(GK_THREADS)
global_solve_kernel(const double* __restrict__ G, int k, int mode, double infl,
                    double* __restrict__ Wout, double* __restrict__ Ainv_out, int* fail_out) {
  extern __shared__ double sm[];
  const int ks = k | 1;
  double* M = sm;
  double* M2 = M + (size_t)k * ks;
  double* F = M2 + (size_t)k * ks;   // k x k (EnKF) or k (ETKF)
  double* wa = F + (size_t)k * k;
  __shared__ int fail;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const double km1 = (double)(k - 1);
  if (tid == 0) fail = 0;
  for (int e = tid; e < k * k; e += nt) {
    const int a = e / k, b = e - a * k;
    M[a * ks + b] = G[e] + (a == b ? km1 : 0.0);
  }
  const int nf = (mode == 0) ? k : k * k;
  for (int e = tid; e < nf; e += nt) F[e] = G[k * k + e];
  __syncthreads();
  lk_cholesky(M, k, ks, &fail);
  lk_tri_inverse(M, M2, k, ks);
  // Pa = Linv^T Linv into M
  for (int e = tid; e < k * k; e += nt) {
    const int a = e / k, b = e - a * k;
    double s = 0.0;
    for (int t = max(a, b); t < k; ++t) s += M2[t * ks + a] * M2[t * ks + b];
    M[a * ks + b] = s;
  }
  __syncthreads();
  if (mode == 1) {
    const double sqi = sqrt(infl);
    for (int e = tid; e < k * k; e += nt) {
      const int a = e / k, b = e - a * k;
      double s = 0.0;
      for (int t = 0; t < k; ++t) s += M[a * ks + t] * F[t * k + b];
      Wout[e] = sqi * ((a == b ? 1.0 : 0.0) + s);
      if (Ainv_out) Ainv_out[e] = M[a * ks + b];
    }
  } else {
    for (int a = warp; a < k; a += nw) {
      double s = 0.0;
      for (int b = lane; b < k; b += 32) s += M[a * ks + b] * F[b];
      s = warp_sum(s);
      if (lane == 0) wa[a] = s;
    }
    __syncthreads();
    lk_cholesky(M, k, ks, &fail);
    const double sq = sqrt(km1);
    for (int e = tid; e < k * k; e += nt) {
      const int j = e / k, i = e - j * k;
      Wout[e] = infl * (wa[j] + (i <= j ? sq * M[j * ks + i] : 0.0));
    }
  }
  __syncthreads();
  if (tid == 0 && fail) *fail_out = 1;
}

// Xa[pt][i] = mean[pt] + sum_j (X[pt][j] - mean[pt]) W[j][i], streamed over all points.
// CTA tile: GA_TP points; W resident in shared memory; 4x4 register tiles.
// Also accumulates sum (x'_j)^2 (background) and sum (xa_i - mean(xa))^2 (analysis) per block.
#define GA_TP 64
__global__ void __launch_bounds__(GK_THREADS)
global_apply_kernel(double* __restrict__ X, const double* __restrict__ mean,
                    const double* __restrict__ W, int64_t npoints, int k,
                    double* __restrict__ spread_partial /*[grid][2] or null*/) {
  extern __shared__ double sm[];
  const int ks = k | 1;
  double* Ws = sm;                          // [k][ks]
  double* Xs = Ws + (size_t)k * ks;         // [GA_TP][ks]   perturbations
  double* Os = Xs + (size_t)GA_TP * ks;     // [GA_TP][ks]   outputs
  double* ms = Os + (size_t)GA_TP * ks;     // [GA_TP]
  __shared__ double red[2][GK_THREADS / 32];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < k * k; e += nt) Ws[(e / k) * ks + (e % k)] = W[e];
  double acc_b = 0.0, acc_a = 0.0;
  const int k4 = (k + 3) / 4;
  for (int64_t base = (int64_t)blockIdx.x * GA_TP; base < npoints; base += (int64_t)gridDim.x * GA_TP) {
    const int np = (int)min((int64_t)GA_TP, npoints - base);
    __syncthreads();
    for (int e = tid; e < np; e += nt) ms[e] = mean[base + e];
    __syncthreads();
    for (int e = tid; e < np * k; e += nt) {
      const int p = e / k, j = e - p * k;
      const double v = X[(base + p) * k + j] - ms[p];
      Xs[p * ks + j] = v;
      acc_b += v * v;
    }
    __syncthreads();
    const int ntile = ((np + 3) / 4) * k4;
    for (int t = tid; t < ntile; t += nt) {
      const int pq = t / k4, iq = t - pq * k4;
      const int p0 = pq * 4, i0 = iq * 4;
      double o[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) o[a][b] = 0.0;
      for (int j = 0; j < k; ++j) {
        double xv[4], wv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xv[a] = (p0 + a < np) ? Xs[(p0 + a) * ks + j] : 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) wv[b] = (i0 + b < k) ? Ws[j * ks + i0 + b] : 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) o[a][b] += xv[a] * wv[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (p0 + a < np && i0 + b < k) Os[(p0 + a) * ks + i0 + b] = ms[p0 + a] + o[a][b];
    }
    __syncthreads();
    for (int e = tid; e < np * k; e += nt) {
      const int p = e / k, j = e - p * k;
      X[(base + p) * k + j] = Os[p * ks + j];
    }
    if (spread_partial) {
      for (int p = warp; p < np; p += nt / 32) {
        double s = 0.0;
        for (int j = lane; j < k; j += 32) s += Os[p * ks + j];
        s = warp_sum(s) * (1.0 / (double)k);
        double q = 0.0;
        for (int j = lane; j < k; j += 32) { double dd = Os[p * ks + j] - s; q += dd * dd; }
        acc_a += q;
      }
    }
  }
  if (spread_partial) {
    acc_b = warp_sum(acc_b);
    acc_a = warp_sum(acc_a);
    if (lane == 0) { red[0][warp] = acc_b; red[1][warp] = acc_a; }
    __syncthreads();
    if (tid == 0) {
      double b = 0.0, a = 0.0;
      for (int w = 0; w < nt / 32; ++w) { b += red[0][w]; a += red[1][w]; }
      spread_partial[2 * blockIdx.x + 0] = b;
      spread_partial[2 * blockIdx.x + 1] = a;
    }
  }
}

// M[o][j] = rinv_o * sum_l Ainv[j][l] Yp[o][l]     (K = X' M^T, see header comment)
__global__ void enkf_gain_factor_kernel(const double* __restrict__ Yp, const double* __restrict__ Ainv,
                                        const double* __restrict__ err, const uint8_t* __restrict__ valid,
                                        int64_t P, int k, double* __restrict__ Mo) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P * k;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = e / k;
    const int j = (int)(e - o * k);
    const double ee = err[o];
    const double rinv = valid[o] ? 1.0 / (ee * ee) : 0.0;
    double s = 0.0;
    for (int l = 0; l < k; ++l) s += Ainv[j * k + l] * Yp[o * k + l];
    Mo[e] = rinv * s;
  }
}

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) <= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

// max/min over K[pt][o] = scale * sum_j (X[pt][j] - mean[pt]) M[o][j]  (EnKF.hpp:199-203): a dense
// n x P x k FP64 contraction (4e10 FMA at C2) whose result is never written -- FP64 tensor path.
// CTA tile 128 points x 128 obs; both operands staged in shared memory as [row][kp] with stride
// == 4 (mod 8) (conflict-free DMMA fragments: A[m][kk] = Xp[m][kk], B[kk][n] = M[n][kk], both the
// 8-rows-x-4-doubles pattern); 8 warps = 2 (points) x 4 (obs), each 64 x 32 = 8 x 4 DMMA tiles:
// 12 fragment loads per 32 MMAs.  Max/min reduced in registers, one atomic pair per warp.
#define GM_TP 128
#define GM_TO 128
#define GM_KC 64
__global__ void __launch_bounds__(GK_THREADS)
enkf_gain_minmax_kernel(const double* __restrict__ X, const double* __restrict__ mean,
                        const double* __restrict__ Mo, int64_t npoints, int64_t P, int k,
                        double scale, double* __restrict__ mm /*[0]=max [1]=min*/) {
  extern __shared__ double sm[];
  const int kp = (k + 3) & ~3;
  const int kc = min(kp, GM_KC), ks = ((kc + 7) & ~7) + 4;     // members staged GM_KC at a time
  double* Xs = sm;                           // [GM_TP][ks]
  double* Ms = Xs + (size_t)GM_TP * ks;      // [GM_TO][ks]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t pt0 = (int64_t)blockIdx.x * GM_TP, ob0 = (int64_t)blockIdx.y * GM_TO;
  const int wp = warp >> 2, wo = warp & 3;       // warp tile: points [64 wp, +64), obs [32 wo, +32)
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const double* xa = Xs + (64 * wp + g) * ks + t;
  const double* mb = Ms + (32 * wo + g) * ks + t;
  for (int k0 = 0; k0 < kp; k0 += kc) {
    const int kn = min(kc, kp - k0);
    __syncthreads();
    for (int e = tid; e < GM_TP * kn; e += GK_THREADS) {
      const int r = e / kn, j = e - r * kn;
      Xs[r * ks + j] = (pt0 + r < npoints && k0 + j < k) ? (X[(pt0 + r) * k + k0 + j] - mean[pt0 + r]) * scale : 0.0;
    }
    for (int e = tid; e < GM_TO * kn; e += GK_THREADS) {
      const int r = e / kn, j = e - r * kn;
      Ms[r * ks + j] = (ob0 + r < P && k0 + j < k) ? Mo[(ob0 + r) * k + k0 + j] : 0.0;
    }
    __syncthreads();
    for (int kk = 0; kk < kn; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = xa[i * 8 * ks + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = mb[j * 8 * ks + kk];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
  }
  double vmax = -INFINITY, vmin = INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t pt = pt0 + 64 * wp + 8 * i + g, ob = ob0 + 32 * wo + 8 * j + 2 * t;
      if (pt < npoints) {
        if (ob < P) { vmax = fmax(vmax, acc[i][j][0]); vmin = fmin(vmin, acc[i][j][0]); }
        if (ob + 1 < P) { vmax = fmax(vmax, acc[i][j][1]); vmin = fmin(vmin, acc[i][j][1]); }
      }
    }
  for (int off = 16; off; off >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
  }
  if (lane == 0) {
    atomic_max_double(mm + 0, vmax);
    atomic_min_double(mm + 1, vmin);
  }
}

// sum d^2 and min/max variance over valid obs (one block)
__global__ void obs_scalar_stats_kernel(const double* __restrict__ d, const double* __restrict__ err,
                                        const uint8_t* __restrict__ valid, int64_t P,
                                        double* __restrict__ out /*[0]=sum d^2 [1]=min var [2]=max var*/) {
  __shared__ double r0[32], r1[32], r2[32];
  double s = 0.0, vmin = INFINITY, vmax = 0.0;
  for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
    s += d[i] * d[i];
    if (valid[i]) { double v = err[i] * err[i]; vmin = fmin(vmin, v); vmax = fmax(vmax, v); }
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { r0[warp] = s; r1[warp] = vmin; r2[warp] = vmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = INFINITY, c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r0[w]; b = fmin(b, r1[w]); c = fmax(c, r2[w]); }
    out[0] = a; out[1] = b; out[2] = c;
  }
}

// all eigenvalues of the k x k symmetric PSD matrix C by one-CTA Jacobi (one-sided solver applied
// to C itself: columns converge to lambda_i u_i, so |column| = lambda_i)
__global__ void __launch_bounds__(GK_THREADS)
sym_eigvals_kernel(const double* __restrict__ Cg, int k, double* __restrict__ out /*[k]*/) {
  extern __shared__ double sm[];
  const int ks = k | 1;
  double* M = sm;
  __shared__ unsigned long long s_maxrel;
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) M[(e / k) * ks + (e % k)] = Cg[e];
  __syncthreads();
  lk_jacobi<4>(M, k, ks, 60, 1e-12, &s_maxrel);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = warp; c < k; c += nw) {
    double s2 = 0.0;
    for (int r = lane; r < k; r += 32) s2 += M[r * ks + c] * M[r * ks + c];
    s2 = warp_sum(s2);
    if (lane == 0) out[c] = sqrt(s2);
  }
}
