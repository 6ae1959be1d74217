// letkf_ns.cuh -- CANONICAL column kernel with a GEMM-only symmetric square root, full products.
//
// The canonical LETKF needs W = sqrt(k-1) A^{-1/2} and w = A^{-1} g for the SPD matrix
// A = (k-1)/infl I + C.  The symmetric square root is unique, so any route to A^{-1/2} gives the
// same transform as the eigen-decomposition W = U sqrt((k-1)/L) U^T (letkf_v2.cuh), to rounding.
// Here it is the coupled Newton-Schulz iteration (Higham, Functions of Matrices, eq. 6.35)
//     T = (3I - Z Y)/2,  Y <- Y T,  Z <- T Z;   Y -> A^{1/2}, Z -> A^{-1/2}
// started from Z0 = q(A), Y0 = A Z0 with q a Chebyshev approximation of x^{-1/2} on the spectral
// interval [lmin, lmin + ||C||_F], lmin = (k-1)/infl the exact lower end (ns_chebyshev_start).
// Every iterate is a polynomial in A; ~7 iterations x 3 products for the C5 conditioning, each a
// real dense k x k x k FP64 contraction out of shared memory -- so they run on the FP64 tensor
// path (DMMA, mma.sync.m8n8k4.f64): an FMA-pipe version with 5 x 5 register tiles is shared-memory
// bound on B200 (4 (TM+TN)/(TM TN) = 1.6 wavefront-cycles per FP64-pipe cycle, measured 65 % smem
// vs 50 % FP64 utilisation), whereas DMMA fragments need ~1.3 eight-byte loads per 256-FMA tile.
// No rotations, shuffles or rsqrt chains, and the update needs ONE product (X' W) instead of two.
//
// This kernel computes every product in full (all tiles, true products P Q), which is stable for any
// conditioning: the identities Z f(YZ) = f(ZY) Z hold structurally, commutativity is not assumed.
// It needs four padded k x k buffers (one CTA per SM, 24 <= k <= 80) and is the fallback of the
// packed symmetric kernel (letkf_nsp.cuh), which is ~2x cheaper per product and runs two columns
// per SM but is only accurate while cond(A) is moderate (NS_SYM_COND_MAX).
#pragma once
#include "letkf_kernels.cuh"

#define NS_THREADS 256   /* 512 (16 warps, 128 regs) measured 10 % slower: C5q 69.8 vs 63.0 ms */
#define NS_WARPS (NS_THREADS / 32)
#define NS_PCH 32
#define NS_SELCAP 512
#define NS_MAX_ITERS 40
// Symmetric-tile products (upper triangle computed, lower implied) assume the iterates commute;
// rounding breaks that and the defect grows with cond(A).  Error of W against the eigen-decomposition
// (numpy emulation of both variants on the test columns, tools notes in DESIGN.md):
//   bound < 100: 1e-14 | 100-200: 1e-13 | 200-300: 5e-13 | 300-500: 3e-12 | 700-1000: 3e-10 | 1e4: 5e-8
// with bound = (lmin + ||C||_F) / lmin >= cond(A); full products stay at 1e-14 .. 1e-13 throughout.
#define NS_SYM_COND_MAX 256.0

// Matrices are padded to kp = 8 ceil(k/8) rows/cols (DMMA tiles) with row stride ks == 4 (mod 8)
// doubles: both fragment patterns -- A: 8 rows x 4 consecutive doubles, B: 4 rows x 8 consecutive
// doubles -- then touch every bank exactly twice (256 B in the minimum 2 wavefronts), and rows
// stay 16-byte aligned for the 128-bit accumulator stores.
__host__ __device__ inline int ns_kp(int k) { return (k + 7) & ~7; }
__host__ __device__ inline int ns_stride(int k) { return ns_kp(k) + 4; }

__device__ __forceinline__ double block_reduce(double v, bool is_max, double* red /*[NS_WARPS]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < NS_THREADS / 32; ++w) r = is_max ? fmax(r, red[w]) : r + red[w];
  return r;
}

// (An FMA-pipe micro-kernel with exact TM x TM register tiles and 128-bit shared loads preceded the
// DMMA products below; it measured shared-memory bound -- 65 % smem wavefronts vs 50 % FP64 pipe,
// profiles/r01_ncu_column_kernel.json entry b -- and was removed.  The products must be TRUE
// products P Q: P^T Q, equal for exactly symmetric iterates, lets rounding asymmetry grow like
// cond(A) per step and the coupled iteration diverges for cond ~ 1e5.)

// ---- upper-triangular tile bookkeeping for the SYRK accumulation of C (phase 1): each warp takes a
// contiguous chunk of the row-major upper-triangle enumeration (k = 80: 55 tiles, 7 per warp), and
// A is written out mirrored.
template <int NTW>
struct NsTiles {
  int ti[NTW], tj[NTW];
  int n;
};
template <int NTW>
__device__ __forceinline__ NsTiles<NTW> ns_tiles(int kp, int warp) {
  const int nt = kp >> 3, E = nt * (nt + 1) / 2;
  const int e0 = (E * warp) / (NS_THREADS / 32), e1 = (E * (warp + 1)) / (NS_THREADS / 32);
  NsTiles<NTW> w;
  w.n = e1 - e0;
  int i = 0, rowstart = 0;              // locate e0: row i starts at rowstart and has nt - i tiles
  while (e0 >= rowstart + (nt - i)) { rowstart += nt - i; ++i; }
  int j = i + (e0 - rowstart);
#pragma unroll
  for (int n = 0; n < NTW; ++n) {
    w.ti[n] = i; w.tj[n] = j;
    if (n + 1 < w.n) { if (++j == nt) { ++i; j = i; } }
  }
  return w;
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// f(row, col, v0, v1, offdiag): accumulator pair at (row, col), (row, col + 1); offdiag tiles must
// also be written transposed by the caller
template <int NTW, typename F>
__device__ __forceinline__ void ns_foreach_sym(const NsTiles<NTW>& w, int lane, double (&acc)[NTW][2], F&& f) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int n = 0; n < NTW; ++n)
    if (n < w.n) f(w.ti[n] * 8 + g, w.tj[n] * 8 + 2 * t, acc[n][0], acc[n][1], w.ti[n] != w.tj[n]);
}

__device__ __forceinline__ void ns_store_sym(double* dst, int ks, int i, int j, double v0, double v1, bool offdiag) {
  *reinterpret_cast<double2*>(dst + i * ks + j) = make_double2(v0, v1);
  if (offdiag) { dst[j * ks + i] = v0; dst[(j + 1) * ks + i] = v1; }
}

// ---- FP64 tensor-core product, all tiles (used when cond(A) is large: see ns_iterate).
// Warp w = (rh, cq): rh = w / 4 picks a half of the tile rows, cq = w % 4 a quarter of the tile
// columns; the quarters' sizes are listed in opposite order for the two halves so that each SM
// sub-partition (warps w and w + 4) gets the same number of MMAs (k = 80: 10 x 10 tiles -> warps
// of 5 x 3 and 5 x 2 tiles, 25 MMAs per k-step per sub-partition).
struct NsWarpTile { int tr0, ntr, tc0, ntc; };
__device__ __forceinline__ NsWarpTile ns_warp_tile(int kp, int warp) {
  // NS_WARPS / 4 row groups x 4 column quarters; odd row groups list the quarter sizes in reverse
  // order, so the four warps of one SM sub-partition (w, w+4, w+8, ...) carry equal MMA counts
  constexpr int RG = NS_WARPS / 4;
  const int nt = kp >> 3, rg = warp >> 2, cq = warp & 3;
  NsWarpTile w;
  w.tr0 = (nt * rg) / RG;
  w.ntr = (nt * (rg + 1)) / RG - w.tr0;
  const int base = nt >> 2, rem = nt & 3;
  int start = 0;
  w.tc0 = 0; w.ntc = 0;
  for (int i = 0; i < 4; ++i) {
    const int qi = (rg & 1) ? 3 - i : i;
    const int sz = base + (qi < rem ? 1 : 0);
    if (i == cq) { w.tc0 = start; w.ntc = sz; }
    start += sz;
  }
  return w;
}

template <int RT, int CT>
__device__ __forceinline__ void ns_mm_full(const double* __restrict__ Pm, const double* __restrict__ Qm,
                                           int kp, int ks, const NsWarpTile& w, int lane,
                                           double (&acc)[RT][CT][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const double* pa = Pm + ((w.tr0 * 8 + g) * ks + t);
  const double* qb = Qm + (t * ks + w.tc0 * 8 + g);
#pragma unroll 2
  for (int kk = 0; kk < kp; kk += 4) {
    double a[RT], b[CT];
#pragma unroll
    for (int i = 0; i < RT; ++i) a[i] = (i < w.ntr) ? pa[i * 8 * ks + kk] : 0.0;
#pragma unroll
    for (int j = 0; j < CT; ++j) b[j] = (j < w.ntc) ? qb[kk * ks + j * 8] : 0.0;
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < CT; ++j)
        if (i < w.ntr && j < w.ntc)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
  }
}

template <int RT, int CT, typename F>
__device__ __forceinline__ void ns_foreach_full(const NsWarpTile& w, int lane, double (&acc)[RT][CT][2], F&& f) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j)
      if (i < w.ntr && j < w.ntc) f((w.tr0 + i) * 8 + g, (w.tc0 + j) * 8 + 2 * t, acc[i][j][0], acc[i][j][1], false);
}

// One full product C = P Q into registers + visitor.
template <int TM>
struct NsProd {
  static constexpr int RT = (2 * TM + NS_WARPS / 4 - 1) / (NS_WARPS / 4), CT = (2 * TM + 3) / 4;
  NsWarpTile ft;
  double facc[RT][CT][2];
  __device__ __forceinline__ void init(int kp, int warp) { ft = ns_warp_tile(kp, warp); }
  __device__ __forceinline__ void mm(const double* Pm, const double* Qm, int kp, int ks, int lane) {
    ns_mm_full<RT, CT>(Pm, Qm, kp, ks, ft, lane, facc);
  }
  template <typename F>
  __device__ __forceinline__ void foreach(int lane, F&& f) { ns_foreach_full<RT, CT>(ft, lane, facc, f); }
};

// Starting point of the iteration: Z0 = q(A) with q the degree-2 Chebyshev interpolant of x^{-1/2}
// on [lo, hi] >= spectrum(A), rescaled so that p(x) = q(x)^2 x -- the spectrum of Z0 Y0 = Z0^2 A --
// is centred on 1 (sampled at 64 points; the iteration only needs |1 - p| < 1, and the exact
// centring only affects the iteration count).  Costs two products (A^2, A Z0) and saves four to five
// iterations of three against the textbook start Z0 = I, Y0 = 2A/(lo + hi).
struct NsStart { double a0, a1, a2; };
__device__ __forceinline__ NsStart ns_chebyshev_start(double lo, double hi) {
  hi = fmax(hi, lo * (1.0 + 1e-6));
  const double t0 = 0.86602540378443865;                   // cos(pi/6); nodes t0, 0, -t0
  const double hw = 0.5 * (hi - lo), mid = 0.5 * (hi + lo);
  const double f0 = rsqrt(mid + hw * t0), f1 = rsqrt(mid), f2 = rsqrt(mid - hw * t0);
  const double c0 = (f0 + f1 + f2) / 3.0;
  const double c1 = (2.0 / 3.0) * t0 * (f0 - f2);
  const double c2 = (2.0 / 3.0) * (0.5 * f0 - f1 + 0.5 * f2);      // T2(t0) = 1/2, T2(0) = -1
  const double m = 1.0 / hw, n = -mid / hw;                         // u = m x + n
  NsStart q;
  q.a0 = c0 + c1 * n + c2 * (2.0 * n * n - 1.0);
  q.a1 = c1 * m + 4.0 * c2 * m * n;
  q.a2 = 2.0 * c2 * m * m;
  // 64 sample points, two per lane (every warp computes the same values; call with all lanes active)
  double pmin = 1e300, pmax = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double x = lo + (hi - lo) * ((double)((threadIdx.x & 31) + 32 * i) * (1.0 / 63.0));
    const double qx = fma(fma(q.a2, x, q.a1), x, q.a0);
    const double p = qx * qx * x;
    pmin = fmin(pmin, p); pmax = fmax(pmax, p);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    pmin = fmin(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
  }
  const double s = sqrt(2.0 / (pmin + pmax));
  q.a0 *= s; q.a1 *= s; q.a2 *= s;
  return q;
}

// Coupled Newton-Schulz  T = (3I - ZY)/2, Y <- YT, Z <- TZ  from a commuting start
// (Z0 = q(A), Y0 = A Z0, A in Tm; see ns_chebyshev_start).  Returns the iterations used, sets ok.
// On exit Zm points at A^{-1/2}.
template <int TM>
__device__ __forceinline__ int ns_iterate(double*& Ym, double*& Zm, double*& Tm, double*& Sm, int kp,
                                          int ks, int warp, int lane, double* red, const NsStart& q,
                                          bool& ok) {
  NsProd<TM> pr;
  pr.init(kp, warp);
  auto store_to = [&](double* dst) {
    pr.foreach(lane, [&](int i, int j, double v0, double v1, bool od) { ns_store_sym(dst, ks, i, j, v0, v1, od); });
  };
  pr.mm(Tm, Tm, kp, ks, lane);                             // A^2
  store_to(Sm);
  __syncthreads();
  for (int e = warp * 32 + lane; e < kp * kp; e += NS_THREADS) {   // Z0 = q(A), exactly symmetric
    const int i = e / kp, j = e - i * kp;
    Zm[i * ks + j] = fma(q.a2, Sm[i * ks + j], fma(q.a1, Tm[i * ks + j], i == j ? q.a0 : 0.0));
  }
  __syncthreads();
  pr.mm(Tm, Zm, kp, ks, lane);                             // Y0 = A Z0
  store_to(Ym);
  __syncthreads();
  int it = 0;
  bool done = false;
  for (; it < NS_MAX_ITERS && !done; ++it) {
    pr.mm(Zm, Ym, kp, ks, lane);                           // Z Y
    double r = 0.0;
    pr.foreach(lane, [&](int i, int j, double v0, double v1, bool od) {
      const double d0 = (i == j ? 1.0 : 0.0), d1 = (i == j + 1 ? 1.0 : 0.0);
      const double e0 = d0 - v0, e1 = d1 - v1;
      r = fmax(r, fmax(fabs(e0), fabs(e1)));
      ns_store_sym(Tm, ks, i, j, d0 + 0.5 * e0, d1 + 0.5 * e1, od);        // (3I - ZY)/2
    });
    r = block_reduce(r, true, red);                         // also publishes T
    done = r < 1e-7;                                         // error after this update ~ r^2
    if (!(r < 1.5)) { ok = false; break; }                   // cannot happen for SPD input; NaN guard
    if (!done) {
      pr.mm(Ym, Tm, kp, ks, lane);                         // Y <- Y T
      store_to(Sm);
    }
    pr.mm(Tm, Zm, kp, ks, lane);                           // Z <- T Z
    __syncthreads();                                         // everyone is done reading Y and Z
    store_to(Ym);
    __syncthreads();
    { double* oldZ = Zm; Zm = Ym; Ym = Sm; Sm = oldZ; }     // Z = new, Y = S, S = old Z
  }
  if (!done) ok = false;
  return it;
}

// Rectangular tile chunk of an (ntr x nt) tile grid for the update product X' Z.
template <int NTA>
__device__ __forceinline__ NsTiles<NTA> ns_rect_tiles(int ntr, int nt, int warp) {
  const int E = ntr * nt;
  const int e0 = (E * warp) / (NS_THREADS / 32), e1 = (E * (warp + 1)) / (NS_THREADS / 32);
  NsTiles<NTA> w;
  w.n = e1 - e0;
#pragma unroll
  for (int n = 0; n < NTA; ++n) {
    const int e = min(e0 + n, E - 1);
    w.ti[n] = e / nt;
    w.tj[n] = e - w.ti[n] * nt;
  }
  return w;
}

template <int TM>
__global__ void __launch_bounds__(NS_THREADS) letkf_ns_kernel(ColParams P, int lch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = NS_THREADS;
  constexpr int NTW = (TM * (2 * TM + 1) + NS_WARPS - 1) / NS_WARPS;   // upper-triangular tiles per warp
  constexpr int NTA = (4 * 2 * TM + NS_WARPS - 1) / NS_WARPS;          // update tiles per warp (lch <= 32 levels)
  const int k = P.k, kp = ns_kp(k), ks = ns_stride(k), nz = P.nz, nt = kp >> 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = NT / 32;
  const int g = lane >> 2, t = lane & 3;
  const size_t msz = (size_t)kp * ks;
  double* Bf[4];
  Bf[0] = reinterpret_cast<double*>(smem_raw);
  Bf[1] = Bf[0] + msz; Bf[2] = Bf[1] + msz; Bf[3] = Bf[2] + msz;
  double* gvec = Bf[3] + msz;
  double* wa = gvec + kp;
  double* tv = wa + kp;
  double* xm = tv + kp;                                        // [lch]
  double* ml = xm + lch;                                       // [lch]
  double* red = ml + lch;                                      // [16]
  double* sel_sq = red + 16;                                   // [NS_SELCAP] sqrt(rho / sigma^2)
  double* sel_d = sel_sq + NS_SELCAP;                          // [NS_SELCAP] sqrt(rho / sigma^2) * d
  int* sel_row = reinterpret_cast<int*>(sel_d + NS_SELCAP);    // [NS_SELCAP] obs row
  int* warp_cnt = sel_row + NS_SELCAP;                         // [32]
  int* s_int = warp_cnt + 32;                                  // [4]
  double* Ych = Bf[2];       // [NS_PCH][ks] staged weighted rows (phase 1; Bf[2..3] idle then)

  const double km1 = (double)(k - 1);
  const bool per_level = P.radius_v > 0.0;
  const int nxf = per_level ? nz : 1;
  const int R = index_reach<false>(P.iv, P.radius);
  const bool redo = P.redo_consume != 0;
  const long long ncols = redo ? (long long)*P.redo_count : (P.cols ? P.ncols : (long long)P.own_nx * P.own_ny);
  const NsTiles<NTW> st = ns_tiles<NTW>(kp, warp);

  for (long long ci = blockIdx.x; ci < ncols; ci += gridDim.x) {
    int lx, ly, lt_b = 0, lt_e = nxf;
    if (redo) {
      const long long item = P.redo_items[ci], c = item / nxf;
      lt_b = (int)(item - c * nxf); lt_e = lt_b + 1;
      lx = (int)(c % P.nx); ly = (int)(c / P.nx);
    } else if (P.cols) { long long c = P.cols[ci]; lx = (int)(c % P.nx); ly = (int)(c / P.nx); }
    else { lx = (int)(ci % P.own_nx); ly = (int)(ci / P.own_nx); }
    int gx = P.gx0 + lx, gy = P.gy0 + ly;
    const long long col = (long long)ly * P.nx + lx;
    index_col_coords<false>(P.iv, col, gx, gy);
    double* Xg = P.X + col * nz * k;
    int col_iters = 0;
    long long col_npl = 0;
    bool col_fail = false;

    for (int lt = lt_b; lt < lt_e; ++lt) {
      // ---------------- 1. selection, gather, C += Yw^T Yw on the FP64 tensor path, g += Yw^T dw
      double cacc[NTW][2];
#pragma unroll
      for (int n = 0; n < NTW; ++n) { cacc[n][0] = 0.0; cacc[n][1] = 0.0; }
      double gacc = 0.0;
      if (tid == 0) s_int[0] = 0;
      __syncthreads();
      int npl = 0;
      int cy0 = 0, cy1 = -1;
      if (P.radius >= 0.0) index_cy_range(P.iv, gy, R, cy0, cy1);
      int cy = cy0, rb = 0, re = 0;
      bool rows_left = (cy <= cy1);
      if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
      while (true) {
        const bool have_batch = rows_left;
        if (have_batch) {
          const int a = rb + tid;
          bool sel = false;
          double sq = 0.0, sd = 0.0;
          int orow = 0;
          if (a < re) {
            double dist;
            sel = index_within<false>(P.iv, col, a, gx, gy, P.radius, &dist);
            double dv = 0.0;
            if (sel && per_level) {
              dv = fabs((double)(P.iv.sz[a] - index_level<false>(P.iv, lt)));
              sel = dv <= P.radius_v;
            }
            if (sel) {
              double rho = 1.0;
              if (P.loc != MDC_LOC_CUTOFF) {
                rho = lk_loc_weight(P.loc, dist, P.radius, P.loc_scale);
                if (per_level) rho *= lk_loc_weight(P.loc, dv, P.radius_v, P.loc_scale_v);
              }
              orow = P.iv.sorted_row[a];
              const double e_ = P.err[orow];
              const double ivar = P.valid[orow] ? 1.0 / (e_ * e_) : 0.0;
              sq = sqrt(rho * (P.use_R ? ivar : 1.0));
              sd = sq > 0.0 ? sq * P.d[orow] : 0.0;            // (weight 0: a NaN missing value must not spread)
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, sel);
          if (lane == 0) warp_cnt[warp] = __popc(bal);
          __syncthreads();
          int off = s_int[0];
          for (int w = 0; w < warp; ++w) off += warp_cnt[w];
          if (sel) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            sel_row[pos] = orow;
            sel_sq[pos] = sq;
            sel_d[pos] = sd;
          }
          __syncthreads();
          if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < nw; ++w) tot += warp_cnt[w];
            s_int[0] += tot;
          }
          rb += NT;
          if (rb >= re) {
            ++cy;
            rows_left = (cy <= cy1);
            if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
          }
          __syncthreads();
        }
        const int nsel = s_int[0];
        if (have_batch && rows_left && nsel <= NS_SELCAP - NT) continue;
        for (int c0 = 0; c0 < nsel; c0 += NS_PCH) {
          const int rows = min(NS_PCH, nsel - c0), rows4 = (rows + 3) & ~3;
          // gather: warp w stages rows w, w + NS_WARPS, ...; all loads issued before the stores
          {
            double v[NS_PCH / NS_WARPS][(2 * TM * 8 + 31) / 32];
#pragma unroll
            for (int q = 0; q < NS_PCH / NS_WARPS; ++q) {
              const int r = warp + NS_WARPS * q;
              const double* src = P.Yp + (long long)sel_row[c0 + min(r, rows - 1)] * k;
#pragma unroll
              for (int jj = 0; jj < (2 * TM * 8 + 31) / 32; ++jj) {
                const int j = lane + 32 * jj;
                v[q][jj] = (r < rows && j < k) ? src[j] : 0.0;
              }
            }
#pragma unroll
            for (int q = 0; q < NS_PCH / NS_WARPS; ++q) {
              const int r = warp + NS_WARPS * q;
              if (r < rows4) {
                const double sq = (r < rows) ? sel_sq[c0 + r] : 0.0;
#pragma unroll
                for (int jj = 0; jj < (2 * TM * 8 + 31) / 32; ++jj) {
                  const int j = lane + 32 * jj;
                  if (j < kp) Ych[r * ks + j] = sq * v[q][jj];
                }
              }
            }
          }
          __syncthreads();
          {
            const double* ya = Ych + t * ks + g;
            for (int kk = 0; kk < rows4; kk += 4) {
              double a[NTW], b[NTW];
#pragma unroll
              for (int n = 0; n < NTW; ++n) {
                if (n == 0 || st.ti[n] != st.ti[n - 1]) a[n] = ya[kk * ks + st.ti[n] * 8]; else a[n] = a[n - 1];
                b[n] = ya[kk * ks + st.tj[n] * 8];
              }
#pragma unroll
              for (int n = 0; n < NTW; ++n)
                if (n < st.n)
                  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                               : "+d"(cacc[n][0]), "+d"(cacc[n][1])
                               : "d"(a[n]), "d"(b[n]));
            }
            if (tid < k) {
              for (int r = 0; r < rows; ++r) gacc = fma(Ych[r * ks + tid], sel_d[c0 + r], gacc);
            }
          }
          __syncthreads();
        }
        npl += nsel;
        if (tid == 0 && nsel) s_int[0] = 0;      // (nsel == 0: already 0, and no barrier since it was read)
        __syncthreads();
        if (!rows_left) break;
      }
      if (lt == lt_b) col_npl = npl;

      // ---------------- 2. Z = A^{-1/2}, A = shift I + C, by coupled Newton-Schulz
      double* Tm = Bf[0];       // A, then T
      double* Zm = Bf[1];
      double* Ym = Bf[2];
      double* Sm = Bf[3];
      bool ok = true;
      if (npl > 0) {
        // spectrum(A) lies in [shift, shift + ||C||_F]: C is PSD, and its smallest eigenvalue is
        // exactly 0 (Y' 1 = 0)
        const double shift = km1 / P.inflation;
        double fro = 0.0;
        ns_foreach_sym<NTW>(st, lane, cacc, [&](int i, int j, double v0, double v1, bool od) {
          if (i < k) {
            const double w = od ? 2.0 : 1.0;
            if (j < k) fro = fma(w * v0, v0, fro);
            if (j + 1 < k) fro = fma(w * v1, v1, fro);
          }
        });
        fro = sqrt(block_reduce(fro, false, red));
        ns_foreach_sym<NTW>(st, lane, cacc, [&](int i, int j, double v0, double v1, bool od) {
          const double d0 = (i == j ? shift : 0.0), d1 = (i == j + 1 ? shift : 0.0);
          const double y0 = (i < k && j < k) ? v0 + d0 : d0;           // shift I on the zero padding
          const double y1 = (i < k && j + 1 < k) ? v1 + d1 : d1;
          ns_store_sym(Tm, ks, i, j, y0, y1, od);
        });
        if (tid < k) gvec[tid] = gacc;
        __syncthreads();
        const NsStart q0 = ns_chebyshev_start(shift, shift + fro);
        const int it = ns_iterate<TM>(Ym, Zm, Tm, Sm, kp, ks, warp, lane, red, q0, ok);
        col_iters = max(col_iters, it);
        // w = Z (Z g)
        if (ok) {
          for (int a = warp; a < k; a += nw) {
            double s = 0.0;
            for (int b = lane; b < k; b += 32) s += Zm[a * ks + b] * gvec[b];
            s = warp_sum(s);
            if (lane == 0) tv[a] = s;
          }
          __syncthreads();
          for (int a = warp; a < k; a += nw) {
            double s = 0.0;
            for (int b = lane; b < k; b += 32) s += Zm[a * ks + b] * tv[b];
            s = warp_sum(s);
            if (lane == 0) wa[a] = s;
          }
          __syncthreads();
        }
      }
      const double sW = sqrt(km1);
      if (!ok) col_fail = true;

      if (P.W_out && P.w_col == col && lt == 0) {
        for (int e = tid; e < k * k; e += NT) {
          const int j = e / k, i = e - j * k;
          double vv;
          if (npl == 0) vv = (i == j) ? sqrt(P.inflation) : 0.0;
          else if (!ok) vv = nan("");
          else vv = wa[j] + sW * Zm[j * ks + i];
          P.W_out[e] = vv;
        }
        __syncthreads();
      }

      // ---------------- 3. X_a = xbar + X' w + sW X' Z on the tensor path, level chunks of lch,
      //                     staged in the two matrix buffers that do not hold Z
      double* Xt = (Zm == Bf[0] || Zm == Bf[1]) ? Bf[2] : Bf[0];      // [lch][ks]
      double* To = Xt + (size_t)lch * ks;                              // [lch][k]
      const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
      if (ok) {
        for (int l0 = lev_b; l0 < lev_e; l0 += lch) {
          const int nl = min(lch, lev_e - l0);
          for (int e = tid; e < nl * k; e += NT) {
            const int l = e / k, j = e - l * k;
            Xt[l * ks + j] = Xg[(long long)l0 * k + e];
          }
          if (kp > k) for (int e = tid; e < nl * (kp - k); e += NT) Xt[(e / (kp - k)) * ks + k + e % (kp - k)] = 0.0;
          __syncthreads();
          for (int l = warp; l < nl; l += nw) {
            double s = 0.0;
            for (int j = lane; j < k; j += 32) s += Xt[l * ks + j];
            s = warp_sum(s) / (double)k;
            double m = 0.0;
            for (int j = lane; j < k; j += 32) {
              const double xp = Xt[l * ks + j] - s;
              Xt[l * ks + j] = xp;
              if (npl > 0) m = fma(xp, wa[j], m);
            }
            m = warp_sum(m);
            if (lane == 0) { xm[l] = s; ml[l] = s + m; }
          }
          __syncthreads();
          if (npl == 0) {
            const double f = sqrt(P.inflation);
            for (int e = tid; e < nl * k; e += NT) { const int l = e / k; To[e] = xm[l] + Xt[l * ks + (e - l * k)] * f; }
          } else {
            const int ntr = (nl + 7) >> 3;
            const NsTiles<NTA> at = ns_rect_tiles<NTA>(ntr, nt, warp);
            double uacc[NTA][2];
#pragma unroll
            for (int n = 0; n < NTA; ++n) { uacc[n][0] = 0.0; uacc[n][1] = 0.0; }
            const double* xa = Xt + g * ks + t;
            const double* zb = Zm + t * ks + g;
#pragma unroll 2
            for (int kk = 0; kk < kp; kk += 4) {
              double a[NTA], b[NTA];
#pragma unroll
              for (int n = 0; n < NTA; ++n) {
                if (n == 0 || at.ti[n] != at.ti[n - 1]) a[n] = xa[at.ti[n] * 8 * ks + kk]; else a[n] = a[n - 1];
                b[n] = zb[kk * ks + at.tj[n] * 8];
              }
#pragma unroll
              for (int n = 0; n < NTA; ++n)
                if (n < at.n)
                  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                               : "+d"(uacc[n][0]), "+d"(uacc[n][1])
                               : "d"(a[n]), "d"(b[n]));
            }
#pragma unroll
            for (int n = 0; n < NTA; ++n)
              if (n < at.n) {
                const int l = at.ti[n] * 8 + g, i = at.tj[n] * 8 + 2 * t;
                if (l < nl) {
                  if (i < k) To[l * k + i] = ml[l] + sW * uacc[n][0];
                  if (i + 1 < k) To[l * k + i + 1] = ml[l] + sW * uacc[n][1];
                }
              }
          }
          __syncthreads();
          for (int e = tid; e < nl * k; e += NT) Xg[(long long)l0 * k + e] = To[e];
          if (P.mean_out) {
            for (int l = warp; l < nl; l += nw) {
              double s = 0.0;
              for (int j = lane; j < k; j += 32) s += To[l * k + j];
              s = warp_sum(s);
              if (lane == 0) P.mean_out[col * nz + l0 + l] = s * (1.0 / (double)k);
            }
          }
          __syncthreads();
        }
      }
    }  // lt
    if (tid == 0) {
      if (!redo) {
        atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)col_npl);
        atomicMax(&P.stats[1], col_npl);
        atomicAdd((unsigned long long*)&P.stats[5], 1ull);
      }
      else atomicAdd((unsigned long long*)&P.stats[6], 1ull);
      atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)col_iters);
      atomicMax(&P.stats[3], (long long)col_iters);
      if (col_fail) atomicAdd((unsigned long long*)&P.stats[4], 1ull);
    }
  }
}

static size_t ns_smem_bytes(int k, int lch) {
  const size_t ks = (size_t)ns_stride(k);
  const size_t dbl = 4 * (size_t)ns_kp(k) * ks + 3 * (size_t)ns_kp(k) + 2 * (size_t)lch + 16 + 2 * NS_SELCAP;
  return dbl * 8 + (size_t)NS_SELCAP * 4 + 32 * 4 + 4 * 4 + 16;
}
