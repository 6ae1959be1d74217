// Ensemble store kernels: member-major host layout <-> [col][lev][member] device layout,
// synthetic fill, mean, checksum.  All HBM-streaming.
#pragma once
#include "mdc_internal.cuh"

// stage: [cb][npts] slices of cb members for flattened points [p0, p0+npts) of the host layout
// [lev][y][x]; X: [col][lev][k].  Reads coalesced over points; each thread writes cb contiguous
// doubles (cb = 8 -> two full 32 B sectors).
__global__ void ens_scatter_members_kernel(double* __restrict__ X, const double* __restrict__ stage,
                                           int64_t p0, int64_t npts, int64_t G, int nz, int k,
                                           int m0, int cb) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  int64_t p = p0 + t;
  int64_t lev = p / G, col = p - lev * G;
  double* dst = X + (col * nz + lev) * k + m0;
  for (int c = 0; c < cb; ++c) dst[c] = stage[(int64_t)c * npts + t];
}

__global__ void ens_gather_members_kernel(const double* __restrict__ X, double* __restrict__ stage,
                                          int64_t p0, int64_t npts, int64_t G, int nz, int k,
                                          int m0, int cb) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  int64_t p = p0 + t;
  int64_t lev = p / G, col = p - lev * G;
  const double* src = X + (col * nz + lev) * k + m0;
  for (int c = 0; c < cb; ++c) stage[(int64_t)c * npts + t] = src[c];
}

// ---- synthetic ensemble (see metada_b200/synthetic.py for the host twin) -------------------
__device__ __forceinline__ double syn_wave(int64_t num, int64_t den) {
  double f = __ddiv_rn((double)(num % den), (double)den);
  double a = __dmul_rn(f, __dsub_rn(1.0, f));
  double b = __dsub_rn(1.0, __dmul_rn(2.0, f));
  return __dmul_rn(__dmul_rn(10.392304845413264, a), b);
}
__device__ __forceinline__ double syn_truth(int gi, int gj, int lev, int gnx, int gny) {
  double wx = syn_wave(3ll * gi, gnx);
  double wy = syn_wave(8ll * gj + gny, 4ll * gny);
  double vz = __dadd_rn(1.0, __dmul_rn(0.01, (double)lev));
  return __dmul_rn(__dmul_rn(wx, wy), vz);
}
__device__ __forceinline__ double syn_noise(uint64_t h) {
  int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48);
  double u = __dsub_rn(__ddiv_rn((double)s, 65536.0), 2.0);
  return __dmul_rn(u, 1.7320508075688772);
}

// one thread per (col, lev, member) element, member fastest -> fully coalesced stores
__global__ void ens_fill_synthetic_kernel(double* __restrict__ X, int nx, int ny, int nz, int k,
                                          int gx0, int gy0, int gnx, int gny, uint64_t seed) {
  int64_t total = (int64_t)nx * ny * nz * k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int m = (int)(e % k);
    int64_t r = e / k;
    int lev = (int)(r % nz);
    int64_t col = r / nz;
    int lx = (int)(col % nx), ly = (int)(col / nx);
    int gi = gx0 + lx, gj = gy0 + ly;
    double t = syn_truth(gi, gj, lev, gnx, gny);
    uint64_t idx = ((uint64_t)lev * (uint64_t)gny + (uint64_t)gj) * (uint64_t)gnx + (uint64_t)gi;
    double nz_ = syn_noise(mdc_hash(seed + (uint64_t)m, idx));
    X[e] = __dadd_rn(t, __dmul_rn(0.5, nz_));
  }
}

// Ensemble::RecomputeMean (Ensemble.hpp:105-114): mean = (0 + x_0 + x_1 + ...) * (1/k), summed in
// member order.  A warp stages 32 consecutive points (32*k contiguous doubles) through shared
// memory so global loads are coalesced; each lane then sums its own row in order.
template <int WARPS>
__global__ void ens_mean_kernel(const double* __restrict__ X, double* __restrict__ mean,
                                int64_t npoints, int k) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = k | 1;  // odd stride: conflict-free row walks
  double* tile = sm + (size_t)warp * 32 * ks;
  const double rk = 1.0 / (double)k;
  for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * 32; base < npoints;
       base += (int64_t)gridDim.x * WARPS * 32) {
    int64_t cnt = npoints - base < 32 ? npoints - base : 32;
    const double* src = X + base * k;
    for (int64_t e = lane; e < cnt * k; e += 32) tile[(e / k) * ks + (e % k)] = src[e];
    __syncwarp();
    if (lane < cnt) {
      double s = 0.0;
      for (int m = 0; m < k; ++m) s = __dadd_rn(s, tile[lane * ks + m]);
      mean[base + lane] = __dmul_rn(s, rk);
    }
    __syncwarp();
  }
}

// mean [col][lev] -> host order [lev][y][x]
__global__ void mean_to_host_order_kernel(const double* __restrict__ mean, double* __restrict__ out,
                                          int64_t G, int nz) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G * nz) return;
  int64_t lev = t / G, col = t - lev * G;
  out[t] = mean[col * nz + lev];
}

__global__ void ens_checksum_kernel(const double* __restrict__ X, int64_t total, double* out2) {
  double s = 0.0, s2 = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    double v = X[e];
    s += v;
    s2 += v * v;
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out2, s);
    atomicAdd(out2 + 1, s2);
  }
}

__global__ void flush_l2_kernel(float4* buf, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    buf[e] = make_float4(0.f, 1.f, 2.f, 3.f);
}
