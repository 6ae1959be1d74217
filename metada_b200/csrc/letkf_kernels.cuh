// Per-grid-column local ensemble transform + state update: one CTA per column.
//
// Replaces LETKF<Tag>::updateGridPoint (LETKF.hpp:152-243) in snapshot semantics:
//   query the bucket index (bit-exact `distance <= radius`), gather the column's local Y' rows
//   into shared memory, accumulate C = Y_l'^T R^-1 Y_l' and g = Y_l'^T R^-1 d, then
//     CANONICAL  : A = (k-1)/infl I + C = L L^T (in-place Cholesky); one-sided (Hestenes) Jacobi
//                  with warp-shuffle reductions on the columns of L -> L V = U S, so A = U S^2 U^T;
//                  W = U sqrt((k-1)/S^2) U^T, w = U S^-2 U^T g, applied as X_a = xbar + X'(w 1^T + W)
//                  through the factor G = U S without ever forming W
//     REF_ETKF   : Pa = A^-1 (Cholesky inverse), wa = Pa g, Wa = sqrt(k-1) chol(Pa), matrix update
//                  (ETKF.hpp:150-169 applied per column)
//     REF_COMPAT : R = I, Pa = A^-1 * infl, Wa = sqrt(k-1) chol(Pa), per-member scaling
//                  xa_i = m + x'_i (wa_i + sum_j Wa_ij)   (LETKF.hpp:214-238, asDiagonal form)
//   and the update of the column's nz*k state block in place.
// FP64 throughout; FP64-pipe / shared-memory bound (tcgen05 has no FP64 kind).
#pragma once
#include "index_kernels.cuh"
#include "mdc_internal.cuh"

#define LK_THREADS 256
#define LK_PCH 32       // local-obs rows staged per SYRK chunk
#define LK_SELCAP 512   // selected-candidate list capacity before a flush
#define LK_LCH 16       // levels per update chunk

struct ColParams {
  double* X;
  double* mean_out;  // [col][lev] analysis mean or nullptr
  int nx, ny, nz, k, own_nx, own_ny, gx0, gy0;
  IndexView iv;
  const double* Yp;  // [row][k]
  const double* d;
  const double* err;
  const uint8_t* valid;
  double radius, radius_v, inflation;
  double loc_scale, loc_scale_v;   // length scales of the exp-type localisation functions
  double kappa_max;                // condition bound up to which the packed Newton-Schulz kernel keeps a transform
  int mode, loc, use_R, max_sweeps;
  double jtol;
  long long* stats;  // [0] sum p_loc [1] max p_loc [2] sum sweeps [3] max sweeps [4] failures [5] columns
  double* W_out;     // optional k*k debug output for column w_col
  long long w_col;
  const long long* cols;  // optional explicit (local linear) column list
  long long ncols;
  // redo list: (local linear column) * nxf + transform index, nxf = nz with per-level transforms,
  // else 1.  The packed Newton-Schulz kernel appends the transforms it must not handle; a kernel
  // launched with redo_consume = 1 processes exactly that list (and leaves stats [0], [1], [5] alone).
  long long* redo_items;
  unsigned* redo_count;
  int redo_consume;
  // transforms with few local observations, handed to the observation-space kernel (letkf_smallp.cuh);
  // same item encoding
  long long* small_items;
  unsigned* small_count;
  // per-level analyses: transforms the classifying first pass (letkf_smallp_classify_kernel) left for
  // the packed k-space kernel; with work_consume = 1 that kernel processes exactly this list
  long long* work_items;
  unsigned* work_count;
  int work_consume;
};

__device__ __forceinline__ double lk_gaspari_cohn(double z) {
  z = fabs(z);
  if (z >= 2.0) return 0.0;
  if (z <= 1.0) return (((-0.25 * z + 0.5) * z + 0.625) * z - 5.0 / 3.0) * z * z + 1.0;
  // close to the end of the support the polynomial cancels to a few 1e-16 of either sign; the taper is >= 0 and the
  // kernels take the square root of the weight
  return fmax(((((z / 12.0 - 0.5) * z + 0.625) * z + 5.0 / 3.0) * z - 5.0) * z + 4.0 - 2.0 / (3.0 * z), 0.0);
}

// The reference's localisation functions (LWEnKF.hpp:597-635).  Not inlined: two double-precision
// exp expansions in the selection loop cost the column kernels ~300 bytes of spills under their
// register cap (C5 probe 35 -> 40 ms), for functions that are evaluated once per selected observation.
__device__ __noinline__ double lk_loc_reference(int loc, double r) {
  switch (loc) {
    case MDC_LOC_GAUSSIAN: return exp(-0.5 * r * r);                                               // :601-602
    case MDC_LOC_EXPONENTIAL: return exp(-r);                                                      // :604-605
    case MDC_LOC_REF_GASPARI_COHN:                                                                 // :624-635
      if (r >= 2.0) return 0.0;
      if (r >= 1.0) { const double z = r - 1.0; return ((-0.25 * z + 0.5) * z + 0.625) * z + 0.125; }
      return (((-0.25 * r + 0.5) * r + 0.625) * r - 5.0) * r + 4.0;
    default: return 1.0;
  }
}
// rho(dist): Gaspari-Cohn taper with the given support, or a reference function of dist / scale
__device__ __forceinline__ double lk_loc_weight(int loc, double dist, double support, double scale) {
  if (loc == MDC_LOC_GASPARI_COHN) return lk_gaspari_cohn(dist / (0.5 * support));
  return lk_loc_reference(loc, dist / scale);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// In-place lower Cholesky of the k x k matrix M (row-major, stride ks). Right-looking.
__device__ void lk_cholesky(double* M, int k, int ks, int* fail) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < k; ++j) {
    if (tid == 0) {
      double v = M[j * ks + j];
      if (!(v > 0.0)) { *fail = 1; v = 1.0; }
      M[j * ks + j] = sqrt(v);
    }
    __syncthreads();
    const double dj = M[j * ks + j];
    for (int i = j + 1 + tid; i < k; i += nt) M[i * ks + j] /= dj;
    __syncthreads();
    const int n = k - j - 1;
    for (int e = tid; e < n * n; e += nt) {
      int i = j + 1 + e / n, c = j + 1 + e % n;
      if (c <= i) M[i * ks + c] -= M[i * ks + j] * M[c * ks + j];
    }
    __syncthreads();
  }
}

__device__ void lk_zero_upper(double* M, int k, int ks) {
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    int i = e / k, c = e % k;
    if (c > i) M[i * ks + c] = 0.0;
  }
  __syncthreads();
}

// Linv (into M2) of the lower-triangular L in M: one warp per column, forward substitution.
__device__ void lk_tri_inverse(const double* L, double* Li, int k, int ks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = warp; c < k; c += nw) {
    for (int i = lane; i < c; i += 32) Li[i * ks + c] = 0.0;
    if (lane == 0) Li[c * ks + c] = 1.0 / L[c * ks + c];
    __syncwarp();
    for (int i = c + 1; i < k; ++i) {
      double s = 0.0;
      for (int t = c + lane; t < i; t += 32) s += L[i * ks + t] * Li[t * ks + c];
      s = warp_sum(s);
      if (lane == 0) Li[i * ks + c] = -s / L[i * ks + i];
      __syncwarp();
    }
  }
  __syncthreads();
}

// One-sided Jacobi on the columns of M (k x k, stride ks odd => conflict-free column walks).
// Round-robin tournament ordering: n-1 steps of n/2 disjoint pairs; a warp owns a pair, lanes own
// rows r = lane + 32 t held in registers between the Gram dot products and the rotation.
template <int NR>
__device__ int lk_jacobi(double* M, int k, int ks, int max_sweeps, double tol,
                         unsigned long long* s_maxrel) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n = (k + 1) & ~1;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (threadIdx.x == 0) *s_maxrel = 0ull;
    __syncthreads();
    double wmax = 0.0;
    for (int step = 0; step < n - 1; ++step) {
      for (int pi = warp; pi < n / 2; pi += nw) {
        int a = (pi == 0) ? n - 1 : (step + pi) % (n - 1);
        int b = (pi == 0) ? step : (step - pi + (n - 1)) % (n - 1);
        int p = min(a, b), q = max(a, b);
        if (q >= k) continue;
        double gp[NR], gq[NR];
        double aa = 0.0, bb = 0.0, gg = 0.0;
#pragma unroll
        for (int t = 0; t < NR; ++t) {
          int r = lane + 32 * t;
          gp[t] = (r < k) ? M[r * ks + p] : 0.0;
          gq[t] = (r < k) ? M[r * ks + q] : 0.0;
          aa += gp[t] * gp[t];
          bb += gq[t] * gq[t];
          gg += gp[t] * gq[t];
        }
        aa = warp_sum(aa); bb = warp_sum(bb); gg = warp_sum(gg);
        double lim = sqrt(aa * bb);
        double rel = (lim > 0.0) ? fabs(gg) / lim : 0.0;
        wmax = fmax(wmax, rel);
        if (rel > 1e-15) {
          double zeta = (bb - aa) / (2.0 * gg);
          double t_ = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = rsqrt(1.0 + t_ * t_);
          double s = c * t_;
#pragma unroll
          for (int t = 0; t < NR; ++t) {
            int r = lane + 32 * t;
            if (r < k) {
              M[r * ks + p] = c * gp[t] - s * gq[t];
              M[r * ks + q] = s * gp[t] + c * gq[t];
            }
          }
        }
      }
      __syncthreads();
    }
    if (lane == 0) atomicMax(s_maxrel, (unsigned long long)__double_as_longlong(wmax));
    __syncthreads();
    double mr = __longlong_as_double((long long)*s_maxrel);
    __syncthreads();
    if (mr < tol) { ++sweep; break; }
  }
  return sweep;
}

// EXT: geographic observations (haversine selection on the lat / lon lattice index) and multi-variable states
// (level map), for the REF modes too -- the reference's own LETKF.hpp arithmetic on a WRF-shaped case.
template <int NR, bool EXT = false>
__global__ void __launch_bounds__(LK_THREADS) letkf_column_kernel(ColParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k, ks = k | 1, nz = P.nz;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = LK_THREADS / 32;
  const bool ref = (P.mode != MDC_MODE_CANONICAL);
  // ---- shared memory carve-up
  double* M = reinterpret_cast<double*>(smem_raw);
  double* M2 = M + (size_t)k * ks;
  double* Ych = M2 + (ref ? (size_t)k * ks : 0);
  double* dw = Ych + (size_t)LK_PCH * k;
  double* gvec = dw + LK_PCH;
  double* lam = gvec + k;
  double* tl = lam + k;
  double* Dv = tl + k;
  double* wa = Dv + k;
  double* Xt = wa + k;
  double* T = Xt + (size_t)LK_LCH * k;
  double* xm = T + (size_t)LK_LCH * k;
  double* ml = xm + LK_LCH;
  double* sel_w = ml + LK_LCH;
  int* sel_pos = reinterpret_cast<int*>(sel_w + LK_SELCAP);
  int* warp_cnt = sel_pos + LK_SELCAP;                // [nw]
  int* s_int = warp_cnt + 32;                         // [0] nsel [1] fail
  unsigned long long* s_maxrel = reinterpret_cast<unsigned long long*>(s_int + 4);

  const double km1 = (double)(k - 1);
  const bool per_level = P.radius_v > 0.0;
  const int nxf = per_level ? nz : 1;
  const int R = index_reach<EXT>(P.iv, P.radius);
  const long long ncols = P.cols ? P.ncols : (long long)P.own_nx * P.own_ny;

  for (long long ci = blockIdx.x; ci < ncols; ci += gridDim.x) {
    int lx, ly;
    if (P.cols) { long long c = P.cols[ci]; lx = (int)(c % P.nx); ly = (int)(c / P.nx); }
    else { lx = (int)(ci % P.own_nx); ly = (int)(ci / P.own_nx); }
    int gx = P.gx0 + lx, gy = P.gy0 + ly;
    const long long col = (long long)ly * P.nx + lx;
    index_col_coords<EXT>(P.iv, col, gx, gy);
    double* Xg = P.X + col * nz * k;
    int col_sweeps = 0;
    long long col_npl = 0;

    for (int lt = 0; lt < nxf; ++lt) {
      // ---------------- 1. selection + accumulation of C (into M) and g
      for (int e = tid; e < k * ks; e += LK_THREADS) M[e] = 0.0;
      for (int e = tid; e < k; e += LK_THREADS) gvec[e] = 0.0;
      if (tid == 0) { s_int[0] = 0; s_int[1] = 0; }
      __syncthreads();
      int npl = 0;
      int cy0 = 0, cy1 = -1;
      if (P.radius >= 0.0) index_cy_range(P.iv, gy, R, cy0, cy1);
      int cy = cy0, rb = 0, re = 0;
      bool rows_left = (cy <= cy1);
      if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
      while (true) {
        // candidate batch [rb, min(rb+THREADS, re))
        bool have_batch = rows_left;
        if (have_batch) {
          int a = rb + tid;
          bool sel = false;
          double rho = 1.0;
          if (a < re) {
            double dist;
            sel = index_within<EXT>(P.iv, col, a, gx, gy, P.radius, &dist);
            double dv = 0.0;
            if (sel && per_level) {
              dv = fabs((double)(P.iv.sz[a] - index_level<EXT>(P.iv, lt)));
              sel = dv <= P.radius_v;
            }
            if (sel && P.mode == MDC_MODE_CANONICAL && P.loc != MDC_LOC_CUTOFF) {
              rho = lk_loc_weight(P.loc, dist, P.radius, P.loc_scale);
              if (per_level) rho *= lk_loc_weight(P.loc, dv, P.radius_v, P.loc_scale_v);
            }
          }
          unsigned bal = __ballot_sync(0xffffffffu, sel);
          if (lane == 0) warp_cnt[warp] = __popc(bal);
          __syncthreads();
          int off = s_int[0];
          for (int w = 0; w < warp; ++w) off += warp_cnt[w];
          if (sel) {
            int pos = off + __popc(bal & ((1u << lane) - 1u));
            sel_pos[pos] = a;
            sel_w[pos] = rho;
          }
          __syncthreads();
          if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < nw; ++w) tot += warp_cnt[w];
            s_int[0] += tot;
          }
          rb += LK_THREADS;
          if (rb >= re) {
            ++cy;
            rows_left = (cy <= cy1);
            if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
          }
          __syncthreads();
        }
        const int nsel = s_int[0];
        if (have_batch && rows_left && nsel <= LK_SELCAP - LK_THREADS) continue;
        // ---- flush: stage weighted rows, rank-PCH update of C and g
        for (int c0 = 0; c0 < nsel; c0 += LK_PCH) {
          const int rows = min(LK_PCH, nsel - c0);
          for (int r = warp; r < rows; r += nw) {
            const int pos = sel_pos[c0 + r];
            const int orow = P.iv.sorted_row[pos];
            double rinv = 1.0;
            if (P.mode != MDC_MODE_REF_COMPAT) {
              const double e_ = P.err[orow];
              const double ivar = P.valid[orow] ? 1.0 / (e_ * e_) : 0.0;
              if (P.mode == MDC_MODE_REF_ETKF) rinv = ivar;
              else rinv = sel_w[c0 + r] * (P.use_R ? ivar : 1.0);
            }
            const double sq = sqrt(rinv);
            const double* src = P.Yp + (long long)orow * k;
            for (int j = lane; j < k; j += 32) Ych[r * k + j] = sq * src[j];
            if (lane == 0) dw[r] = rinv > 0.0 ? sq * P.d[orow] : 0.0;   // (weight 0: a NaN missing value must not spread)
          }
          __syncthreads();
          for (int e = tid; e < k * k; e += LK_THREADS) {
            const int a = e / k, b = e - a * k;
            double s = M[a * ks + b];
            for (int r = 0; r < rows; ++r) s += Ych[r * k + a] * Ych[r * k + b];
            M[a * ks + b] = s;
          }
          for (int a = tid; a < k; a += LK_THREADS) {
            double s = gvec[a];
            for (int r = 0; r < rows; ++r) s += Ych[r * k + a] * dw[r];
            gvec[a] = s;
          }
          __syncthreads();
        }
        npl += nsel;
        if (tid == 0 && nsel) s_int[0] = 0;      // (nsel == 0: already 0, and no barrier since it was read)
        __syncthreads();
        if (!rows_left) break;
      }
      if (lt == 0) col_npl = npl;

      // ---------------- 2. transform
      bool have_xform = npl > 0;
      int sweeps = 0;
      if (have_xform) {
        if (!ref) {
          for (int a = tid; a < k; a += LK_THREADS) M[a * ks + a] += km1 / P.inflation;
          __syncthreads();
          lk_cholesky(M, k, ks, &s_int[1]);
          lk_zero_upper(M, k, ks);
          sweeps = lk_jacobi<NR>(M, k, ks, P.max_sweeps, P.jtol, s_maxrel);
          for (int c = warp; c < k; c += nw) {
            double s2 = 0.0, tg = 0.0;
            for (int r = lane; r < k; r += 32) {
              double v = M[r * ks + c];
              s2 += v * v;
              tg += v * gvec[r];
            }
            s2 = warp_sum(s2); tg = warp_sum(tg);
            if (lane == 0) {
              lam[c] = s2;
              tl[c] = tg / (s2 * s2);
              Dv[c] = sqrt(km1 / s2) / s2;
            }
          }
          __syncthreads();
        } else {
          for (int a = tid; a < k; a += LK_THREADS) M[a * ks + a] += km1;
          __syncthreads();
          lk_cholesky(M, k, ks, &s_int[1]);
          lk_tri_inverse(M, M2, k, ks);
          const double f = (P.mode == MDC_MODE_REF_COMPAT) ? P.inflation : 1.0;
          for (int e = tid; e < k * k; e += LK_THREADS) {
            const int a = e / k, b = e - a * k;
            double s = 0.0;
            for (int t = max(a, b); t < k; ++t) s += M2[t * ks + a] * M2[t * ks + b];
            M[a * ks + b] = s * f;
          }
          __syncthreads();
          for (int a = warp; a < k; a += nw) {
            double s = 0.0;
            for (int b = lane; b < k; b += 32) s += M[a * ks + b] * gvec[b];
            s = warp_sum(s);
            if (lane == 0) wa[a] = s;
          }
          __syncthreads();
          lk_cholesky(M, k, ks, &s_int[1]);
          lk_zero_upper(M, k, ks);
          if (P.mode == MDC_MODE_REF_COMPAT) {
            const double sq = sqrt(km1);
            for (int i = warp; i < k; i += nw) {
              double s = 0.0;
              for (int j = lane; j <= i; j += 32) s += M[i * ks + j];
              s = warp_sum(s);
              if (lane == 0) tl[i] = wa[i] + sq * s;   // per-member scale factor s_i
            }
            __syncthreads();
          }
        }
        __syncthreads();                   // (everyone reads the failure flag before thread 0 can reset it for the
        if (s_int[1]) have_xform = false;  //  next transform)  numeric failure: leave the column unchanged
      }
      col_sweeps = max(col_sweeps, sweeps);

      // optional debug dump of W (level 0)
      if (P.W_out && P.w_col == col && lt == 0) {
        const double sqk = sqrt(km1);
        for (int e = tid; e < k * k; e += LK_THREADS) {
          const int j = e / k, i = e - j * k;
          double v = 0.0;
          if (npl == 0) {
            double f = (P.mode == MDC_MODE_REF_ETKF) ? P.inflation : sqrt(P.inflation);
            v = (P.mode == MDC_MODE_REF_COMPAT) ? (j == 0 ? f : 0.0) : (i == j ? f : 0.0);
          } else if (!have_xform) {
            v = nan("");
          } else if (P.mode == MDC_MODE_CANONICAL) {
            double wj = 0.0, s = 0.0;
            for (int c = 0; c < k; ++c) {
              wj += M[j * ks + c] * tl[c];
              s += M[j * ks + c] * Dv[c] * M[i * ks + c];
            }
            v = wj + s;
          } else if (P.mode == MDC_MODE_REF_ETKF) {
            v = wa[j] + sqk * M[j * ks + i];
          } else {
            v = (j == 0) ? tl[i] : 0.0;
          }
          P.W_out[e] = v;
        }
        __syncthreads();
      }

      // ---------------- 3. update the column's levels in place
      const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
      const bool fail = (npl > 0) && !have_xform;
      if (!fail) {
        for (int l0 = lev_b; l0 < lev_e; l0 += LK_LCH) {
          const int nl = min(LK_LCH, lev_e - l0);
          for (int e = tid; e < nl * k; e += LK_THREADS) Xt[e] = Xg[(long long)l0 * k + e];
          __syncthreads();
          for (int l = warp; l < nl; l += nw) {
            double s = 0.0;
            for (int j = lane; j < k; j += 32) s += Xt[l * k + j];
            s = warp_sum(s) / (double)k;
            if (lane == 0) xm[l] = s;
            const double pre = (P.mode == MDC_MODE_REF_ETKF) ? P.inflation : 1.0;
            for (int j = lane; j < k; j += 32) Xt[l * k + j] = (Xt[l * k + j] - s) * pre;
          }
          __syncthreads();
          if (npl == 0) {
            const double f = (P.mode == MDC_MODE_REF_ETKF) ? 1.0 : sqrt(P.inflation);
            for (int e = tid; e < nl * k; e += LK_THREADS) T[e] = xm[e / k] + Xt[e] * f;
          } else if (P.mode == MDC_MODE_CANONICAL) {
            for (int e = tid; e < nl * k; e += LK_THREADS) {
              const int l = e / k, c = e - l * k;
              double s = 0.0;
              for (int j = 0; j < k; ++j) s += Xt[l * k + j] * M[j * ks + c];
              T[e] = s;
            }
            __syncthreads();
            for (int l = warp; l < nl; l += nw) {
              double s = 0.0;
              for (int c = lane; c < k; c += 32) s += T[l * k + c] * tl[c];
              s = warp_sum(s);
              if (lane == 0) ml[l] = s;
            }
            __syncthreads();
            for (int e = tid; e < nl * k; e += LK_THREADS) T[e] *= Dv[e % k];
            __syncthreads();
            for (int e = tid; e < nl * k; e += LK_THREADS) {
              const int l = e / k, i = e - l * k;
              double s = 0.0;
              for (int c = 0; c < k; ++c) s += T[l * k + c] * M[i * ks + c];
              Xt[e] = (xm[l] + ml[l]) + s;   // reuse Xt as output (each thread owns its element)
            }
            __syncthreads();
            for (int e = tid; e < nl * k; e += LK_THREADS) T[e] = Xt[e];
          } else if (P.mode == MDC_MODE_REF_ETKF) {
            const double sqk = sqrt(km1);
            for (int l = warp; l < nl; l += nw) {
              double s = 0.0;
              for (int j = lane; j < k; j += 32) s += Xt[l * k + j] * wa[j];
              s = warp_sum(s);
              if (lane == 0) ml[l] = s;
            }
            __syncthreads();
            for (int e = tid; e < nl * k; e += LK_THREADS) {
              const int l = e / k, i = e - l * k;
              double s = 0.0;
              for (int j = i; j < k; ++j) s += Xt[l * k + j] * M[j * ks + i];
              T[e] = (xm[l] + ml[l]) + sqk * s;
            }
          } else {  // REF_COMPAT
            for (int e = tid; e < nl * k; e += LK_THREADS) T[e] = xm[e / k] + Xt[e] * tl[e % k];
          }
          __syncthreads();
          for (int e = tid; e < nl * k; e += LK_THREADS) Xg[(long long)l0 * k + e] = T[e];
          if (P.mean_out) {
            for (int l = warp; l < nl; l += nw) {
              double s = 0.0;
              for (int j = lane; j < k; j += 32) s += T[l * k + j];
              s = warp_sum(s);
              if (lane == 0) P.mean_out[col * nz + l0 + l] = s * (1.0 / (double)k);
            }
          }
          __syncthreads();
        }
      }
      if (tid == 0 && fail) atomicAdd((unsigned long long*)&P.stats[4], 1ull);
    }  // lt
    if (tid == 0) {
      atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)col_npl);
      atomicMax(&P.stats[1], col_npl);
      atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)col_sweeps);
      atomicMax(&P.stats[3], (long long)col_sweeps);
      atomicAdd((unsigned long long*)&P.stats[5], 1ull);
    }
  }
}

static size_t lk_smem_bytes(int k, int mode) {
  const size_t ks = (size_t)(k | 1);
  size_t dbl = (size_t)k * ks * (mode == MDC_MODE_CANONICAL ? 1 : 2) + (size_t)LK_PCH * k + LK_PCH +
               5 * (size_t)k + 2 * (size_t)LK_LCH * k + 2 * LK_LCH + LK_SELCAP;
  return dbl * 8 + (size_t)LK_SELCAP * 4 + 32 * 4 + 4 * 4 + 16;
}
