// letkf_v2.cuh -- optimised CANONICAL column kernel (the headline path).
//
// Same mathematics as letkf_kernels.cuh (CANONICAL), restructured around the FP64 pipe:
//   * C = Y_l'^T (rho R^-1) Y_l' accumulated in REGISTER tiles (TM x TM per thread on a 16 x 16
//     thread grid) straight from the staged, sqrt-weighted local rows; no shared-memory RMW.
//   * A = (k-1)/infl I + C is symmetric positive definite, so its singular vectors are its
//     eigenvectors: one-sided (Hestenes) Jacobi is applied to the rows of A itself.  On exit
//     row c = lambda_c u_c^T.  No Cholesky, no V accumulation.
//   * BLOCKED ordering, register resident: the k vectors are cut into blocks of 4; an LG-lane
//     sub-warp group loads one PAIR of blocks (8 vectors, RPL elements per lane each) into
//     registers, does the 16 cross rotations (4 rounds x 4 independent pairs; 12 more inside the
//     blocks once per sweep), and stores them back.  A round-robin tournament over the blocks
//     visits every pair once per sweep.  Shared-memory traffic per rotation drops 4x against the
//     pair-at-a-time kernel and every rotation has 4-way ILP.
//   * Gram dot products use a transposing sub-warp shuffle reduction (10 shuffles for 4 values over
//     16 lanes) that leaves pair i's total on the lanes with (lane & 3) == i; those lanes compute
//     that pair's (c, s) with two rsqrt's and broadcast them -- one pass of rotation-parameter
//     arithmetic serves four rotations.  Vector norms are tracked (a' = a - t g, b' = b + t g).
//   * X_a = xbar + X'(w 1^T + W) applied through the factor G = U Lambda without forming W, as two
//     register-tiled (TL x TM) products per level chunk.
#pragma once
#include "letkf_kernels.cuh"

#define V2_PCH 32
#define V2_SELCAP 512

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <int LG>
__device__ __forceinline__ double shg(double v, int src) { return __shfl_sync(0xffffffffu, v, src, LG); }

// sum over the LG-lane group; every lane gets the total
template <int LG>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LG / 2; o; o >>= 1) v += shx(v, o);
  return v;
}

// 4 values per lane -> lane lg ends with the group total of value (lg & 3)
template <int LG>
__device__ __forceinline__ double group_sum4_transposed(double g0, double g1, double g2, double g3, int lg) {
  const bool b0 = lg & 1, b1 = lg & 2;
  const double rx = shx(b0 ? g0 : g1, 1), ry = shx(b0 ? g2 : g3, 1);
  const double k0 = (b0 ? g1 : g0) + rx, k1 = (b0 ? g3 : g2) + ry;
  const double r = shx(b1 ? k0 : k1, 2);
  double gm = (b1 ? k1 : k0) + r;
#pragma unroll
  for (int o = 4; o < LG; o <<= 1) gm += shx(gm, o);
  return gm;
}

__device__ __forceinline__ double sel4(int i, double a, double b, double c, double d) {
  return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d));
}

// One round of 4 independent rotations (P_i, Q_i) on the register-resident vectors v[8][RPL].
template <int LG, int RPL, int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
__device__ __forceinline__ void jr_round(double (&v)[8][RPL], double (&nv)[8], int lg, double tol2,
                                         bool& notconv) {
  double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
  for (int t = 0; t < RPL; ++t) {
    g0 = fma(v[P0][t], v[Q0][t], g0);
    g1 = fma(v[P1][t], v[Q1][t], g1);
    g2 = fma(v[P2][t], v[Q2][t], g2);
    g3 = fma(v[P3][t], v[Q3][t], g3);
  }
  const double gm = group_sum4_transposed<LG>(g0, g1, g2, g3, lg);
  const int mi = lg & 3;
  const double al = sel4(mi, nv[P0], nv[P1], nv[P2], nv[P3]);
  const double be = sel4(mi, nv[Q0], nv[Q1], nv[Q2], nv[Q3]);
  const double gg = gm * gm, ab = al * be;
  notconv |= gg > tol2 * ab;
  double c = 1.0, s = 0.0, tg = 0.0;
  const bool rot = gg > 1e-30 * ab;
  if (!__any_sync(0xffffffffu, rot)) return;   // every pair of this warp already orthogonal
  if (rot) {
    // tan(2 theta) = 2 g / (be - al), |theta| <= pi/4:
    //   cos(2 theta) = |d| / h, c = sqrt((1 + cos 2theta)/2), s = sign(d) g / (h c), t = s / c
    const double dl = be - al;
    const double rh = rsqrt(fma(dl, dl, 4.0 * gg));
    const double u = fma(0.5 * fabs(dl), rh, 0.5);
    const double rc = rsqrt(u);
    c = u * rc;
    s = copysign(gm * rh * rc, dl * gm);
    if (dl == 0.0) s = copysign(gm * rh * rc, gm);
    tg = s * rc * gm;
  }
  const double c0 = shg<LG>(c, 0), c1 = shg<LG>(c, 1), c2 = shg<LG>(c, 2), c3 = shg<LG>(c, 3);
  const double s0 = shg<LG>(s, 0), s1 = shg<LG>(s, 1), s2 = shg<LG>(s, 2), s3 = shg<LG>(s, 3);
  const double t0 = shg<LG>(tg, 0), t1 = shg<LG>(tg, 1), t2 = shg<LG>(tg, 2), t3 = shg<LG>(tg, 3);
#pragma unroll
  for (int t = 0; t < RPL; ++t) {
    double p, q;
    p = v[P0][t]; q = v[Q0][t]; v[P0][t] = fma(c0, p, -s0 * q); v[Q0][t] = fma(s0, p, c0 * q);
    p = v[P1][t]; q = v[Q1][t]; v[P1][t] = fma(c1, p, -s1 * q); v[Q1][t] = fma(s1, p, c1 * q);
    p = v[P2][t]; q = v[Q2][t]; v[P2][t] = fma(c2, p, -s2 * q); v[Q2][t] = fma(s2, p, c2 * q);
    p = v[P3][t]; q = v[Q3][t]; v[P3][t] = fma(c3, p, -s3 * q); v[Q3][t] = fma(s3, p, c3 * q);
  }
  nv[P0] -= t0; nv[Q0] += t0;
  nv[P1] -= t1; nv[Q1] += t1;
  nv[P2] -= t2; nv[Q2] += t2;
  nv[P3] -= t3; nv[Q3] += t3;
}

// Blocked one-sided Jacobi on the k vectors M[c][0..k) (stride ks).  Returns sweeps used.
template <int NT, int LG, int RPL>
__device__ int jacobi_blocked(double* __restrict__ M, int k, int ks, int max_sweeps, double tol,
                              int* s_flag) {
  const int tid = threadIdx.x;
  const int grp = tid / LG, lg = tid % LG;
  constexpr int GROUPS = NT / LG;
  const int warp_first_grp = (tid & ~31) / LG;   // warp-uniform: skip warps with no active group
  const int nb = ((k + 7) / 8) * 2;       // blocks of 4 vectors, even count
  const int npairs = nb / 2;
  const int npairs_pad = ((npairs + GROUPS - 1) / GROUPS) * GROUPS;
  const double tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    bool notconv = false;
    for (int bs = 0; bs < nb - 1; ++bs) {
      for (int pi = grp; pi < npairs_pad; pi += GROUPS) {
        if (pi - grp + warp_first_grp >= npairs) continue;   // whole warp idle this pass
        const bool active = pi < npairs;
        int bi = 0, bj = 0;
        if (active) {
          const int a = (pi == 0) ? nb - 1 : (bs + pi) % (nb - 1);
          const int b = (pi == 0) ? bs : (bs - pi + (nb - 1)) % (nb - 1);
          bi = min(a, b); bj = max(a, b);
        }
        double v[8][RPL];
        double nv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = (i < 4 ? bi * 4 + i : bj * 4 + (i - 4));
          double n2 = 0.0;
#pragma unroll
          for (int t = 0; t < RPL; ++t) {
            const int r = lg + LG * t;
            const double x = (active && c < k && r < k) ? M[c * ks + r] : 0.0;
            v[i][t] = x;
            n2 = fma(x, x, n2);
          }
          nv[i] = group_sum<LG>(n2);
        }
        if (bs == 0) {  // pairs inside the two blocks, once per sweep
          jr_round<LG, RPL, 0, 1, 2, 3, 4, 5, 6, 7>(v, nv, lg, tol2, notconv);
          jr_round<LG, RPL, 0, 2, 1, 3, 4, 6, 5, 7>(v, nv, lg, tol2, notconv);
          jr_round<LG, RPL, 0, 3, 1, 2, 4, 7, 5, 6>(v, nv, lg, tol2, notconv);
        }
        jr_round<LG, RPL, 0, 4, 1, 5, 2, 6, 3, 7>(v, nv, lg, tol2, notconv);
        jr_round<LG, RPL, 0, 5, 1, 6, 2, 7, 3, 4>(v, nv, lg, tol2, notconv);
        jr_round<LG, RPL, 0, 6, 1, 7, 2, 4, 3, 5>(v, nv, lg, tol2, notconv);
        jr_round<LG, RPL, 0, 7, 1, 4, 2, 5, 3, 6>(v, nv, lg, tol2, notconv);
        if (active) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = (i < 4 ? bi * 4 + i : bj * 4 + (i - 4));
            if (c < k) {
#pragma unroll
              for (int t = 0; t < RPL; ++t) {
                const int r = lg + LG * t;
                if (r < k) M[c * ks + r] = v[i][t];
              }
            }
          }
        }
      }
      __syncthreads();
    }
    if (notconv) *s_flag = 1;
    __syncthreads();
    const int f = *s_flag;
    __syncthreads();
    if (!f) { ++sweep; break; }
  }
  return sweep;
}

// Thread grid (NT/16) x 16.  TM = ceil(k/16): tile width (members) of SYRK and of the update;
// TMY = ceil(k/(NT/16)): tile height of SYRK.  MINB: CTAs per SM the register budget is cut for.
template <int NT, int MINB, int LG, int RPL, int TM, int TMY, bool EXT = false>
__global__ void __launch_bounds__(NT, MINB) letkf_canonical_kernel(ColParams P, int lch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TL = 3;  // levels per thread tile in the update
  constexpr int NTY = NT / 16;
  const int k = P.k, ks = k | 1, nz = P.nz;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = NT / 32;
  double* M = reinterpret_cast<double*>(smem_raw);             // [k][ks]
  double* U = M + (size_t)k * ks;                              // union: Ych|dw  or  Xt|T
  double* Ych = U;                                             // [V2_PCH][k]
  double* dw = Ych + (size_t)V2_PCH * k;                       // [V2_PCH]
  double* Xt = U;                                              // [lch][k]
  double* T = Xt + (size_t)lch * k;                            // [lch][k]
  const size_t usz = max((size_t)V2_PCH * k + V2_PCH, (size_t)2 * lch * k);
  double* gvec = U + usz;
  double* tl = gvec + k;
  double* Dv = tl + k;
  double* xm = Dv + k;                                         // [lch]
  double* ml = xm + lch;                                       // [lch]
  double* sel_w = ml + lch;                                    // [V2_SELCAP]
  int* sel_pos = reinterpret_cast<int*>(sel_w + V2_SELCAP);    // [V2_SELCAP]
  int* warp_cnt = sel_pos + V2_SELCAP;                         // [32]
  int* s_int = warp_cnt + 32;                                  // [4]

  const double km1 = (double)(k - 1);
  const bool per_level = P.radius_v > 0.0;
  const int nxf = per_level ? nz : 1;
  const int R = index_reach<EXT>(P.iv, P.radius);
  const bool redo = P.redo_consume != 0;
  const long long ncols = redo ? (long long)*P.redo_count : (P.cols ? P.ncols : (long long)P.own_nx * P.own_ny);
  const int ty = tid >> 4, tx = tid & 15;   // SYRK thread grid

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
    index_col_coords<EXT>(P.iv, col, gx, gy);
    double* Xg = P.X + col * nz * k;
    int col_sweeps = 0;
    long long col_npl = 0;

    for (int lt = lt_b; lt < lt_e; ++lt) {
      // ---------------- 1. selection, gather, register-tiled C += Yw^T Yw, g += Yw^T dw
      double acc[TMY][TM];
#pragma unroll
      for (int a = 0; a < TMY; ++a)
#pragma unroll
        for (int b = 0; b < TM; ++b) acc[a][b] = 0.0;
      double gacc = 0.0;
      if (tid == 0) { s_int[0] = 0; s_int[1] = 0; }
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
          double rho = 1.0;
          if (a < re) {
            double dist;
            sel = index_within<EXT>(P.iv, col, a, gx, gy, P.radius, &dist);
            double dv = 0.0;
            if (sel && per_level) {
              dv = fabs((double)(P.iv.sz[a] - index_level<EXT>(P.iv, lt)));
              sel = dv <= P.radius_v;
            }
            if (sel && P.loc != MDC_LOC_CUTOFF) {
              rho = lk_loc_weight(P.loc, dist, P.radius, P.loc_scale);
              if (per_level) rho *= lk_loc_weight(P.loc, dv, P.radius_v, P.loc_scale_v);
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, sel);
          if (lane == 0) warp_cnt[warp] = __popc(bal);
          __syncthreads();
          int off = s_int[0];
          for (int w = 0; w < warp; ++w) off += warp_cnt[w];
          if (sel) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            sel_pos[pos] = a;
            sel_w[pos] = rho;
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
        if (have_batch && rows_left && nsel <= V2_SELCAP - NT) continue;
        for (int c0 = 0; c0 < nsel; c0 += V2_PCH) {
          const int rows = min(V2_PCH, nsel - c0);
          for (int r = warp; r < rows; r += nw) {
            const int orow = P.iv.sorted_row[sel_pos[c0 + r]];
            const double e_ = P.err[orow];
            const double ivar = P.valid[orow] ? 1.0 / (e_ * e_) : 0.0;
            const double sq = sqrt(sel_w[c0 + r] * (P.use_R ? ivar : 1.0));
            const double* src = P.Yp + (long long)orow * k;
            for (int j = lane; j < k; j += 32) Ych[r * k + j] = sq * src[j];
            if (lane == 0) dw[r] = sq > 0.0 ? sq * P.d[orow] : 0.0;
          }
          __syncthreads();
          for (int r = 0; r < rows; ++r) {
            double ya[TMY], yb[TM];
#pragma unroll
            for (int a = 0; a < TMY; ++a) { const int ia = ty + NTY * a; ya[a] = ia < k ? Ych[r * k + ia] : 0.0; }
#pragma unroll
            for (int b = 0; b < TM; ++b) { const int ib = tx + 16 * b; yb[b] = ib < k ? Ych[r * k + ib] : 0.0; }
#pragma unroll
            for (int a = 0; a < TMY; ++a)
#pragma unroll
              for (int b = 0; b < TM; ++b) acc[a][b] = fma(ya[a], yb[b], acc[a][b]);
            if (tid < k) gacc = fma(Ych[r * k + tid], dw[r], gacc);
          }
          __syncthreads();
        }
        npl += nsel;
        if (tid == 0 && nsel) s_int[0] = 0;      // (nsel == 0: already 0, and no barrier since it was read)
        __syncthreads();
        if (!rows_left) break;
      }
      if (lt == 0) col_npl = npl;

      // ---------------- 2. eigen-decomposition of A by blocked one-sided Jacobi on its rows
      int sweeps = 0;
      if (npl > 0) {
        const double shift = km1 / P.inflation;
#pragma unroll
        for (int a = 0; a < TMY; ++a)
#pragma unroll
          for (int b = 0; b < TM; ++b) {
            const int ia = ty + NTY * a, ib = tx + 16 * b;
            if (ia < k && ib < k) M[ia * ks + ib] = acc[a][b] + (ia == ib ? shift : 0.0);
          }
        if (tid < k) gvec[tid] = gacc;
        __syncthreads();
        sweeps = jacobi_blocked<NT, LG, RPL>(M, k, ks, P.max_sweeps, P.jtol, &s_int[2]);
        for (int c = warp; c < k; c += nw) {
          double s2 = 0.0, tg = 0.0;
          for (int r = lane; r < k; r += 32) {
            const double x = M[c * ks + r];
            s2 = fma(x, x, s2);
            tg = fma(x, gvec[r], tg);
          }
          s2 = warp_sum(s2); tg = warp_sum(tg);
          if (lane == 0) {
            const double lam = sqrt(s2);               // |row c| = lambda_c
            tl[c] = tg / (s2 * lam);                   // (G_c . g) / lambda^3
            Dv[c] = sqrt(km1 / lam) / s2;              // sqrt((k-1)/lambda) / lambda^2
          }
        }
        __syncthreads();
      }
      col_sweeps = max(col_sweeps, sweeps);

      if (P.W_out && P.w_col == col && lt == 0) {
        for (int e = tid; e < k * k; e += NT) {
          const int j = e / k, i = e - j * k;
          double vv;
          if (npl == 0) vv = (i == j) ? sqrt(P.inflation) : 0.0;
          else {
            double wj = 0.0, s = 0.0;
            for (int c = 0; c < k; ++c) {
              wj += M[c * ks + j] * tl[c];
              s += M[c * ks + j] * Dv[c] * M[c * ks + i];
            }
            vv = wj + s;
          }
          P.W_out[e] = vv;
        }
        __syncthreads();
      }

      // ---------------- 3. X_a = xbar + ((X' G) D) G^T + (X' G) tl, level chunks of lch
      const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
      for (int l0 = lev_b; l0 < lev_e; l0 += lch) {
        const int nl = min(lch, lev_e - l0);
        for (int e = tid; e < nl * k; e += NT) Xt[e] = Xg[(long long)l0 * k + e];
        __syncthreads();
        for (int l = warp; l < nl; l += nw) {
          double s = 0.0;
          for (int j = lane; j < k; j += 32) s += Xt[l * k + j];
          s = warp_sum(s) / (double)k;
          if (lane == 0) xm[l] = s;
          for (int j = lane; j < k; j += 32) Xt[l * k + j] -= s;
        }
        __syncthreads();
        if (npl == 0) {
          const double f = sqrt(P.inflation);
          for (int e = tid; e < nl * k; e += NT) T[e] = xm[e / k] + Xt[e] * f;
          __syncthreads();
        } else {
          const int nlt = (nl + TL - 1) / TL;
          // T[l][c] = sum_j X'[l][j] M[c][j]        (c = tx + 16 b: stride ks across lanes)
          for (int lt2 = ty; lt2 < nlt; lt2 += NTY) {
            double o[TL][TM];
#pragma unroll
            for (int a = 0; a < TL; ++a)
#pragma unroll
              for (int b = 0; b < TM; ++b) o[a][b] = 0.0;
            for (int j = 0; j < k; ++j) {
              double xv[TL], mv[TM];
#pragma unroll
              for (int a = 0; a < TL; ++a) { const int l = lt2 + nlt * a; xv[a] = l < nl ? Xt[l * k + j] : 0.0; }
#pragma unroll
              for (int b = 0; b < TM; ++b) { const int c = tx + 16 * b; mv[b] = c < k ? M[c * ks + j] : 0.0; }
#pragma unroll
              for (int a = 0; a < TL; ++a)
#pragma unroll
                for (int b = 0; b < TM; ++b) o[a][b] = fma(xv[a], mv[b], o[a][b]);
            }
#pragma unroll
            for (int a = 0; a < TL; ++a)
#pragma unroll
              for (int b = 0; b < TM; ++b) {
                const int l = lt2 + nlt * a, c = tx + 16 * b;
                if (l < nl && c < k) T[l * k + c] = o[a][b];
              }
          }
          __syncthreads();
          for (int l = warp; l < nl; l += nw) {
            double s = 0.0;
            for (int c = lane; c < k; c += 32) s += T[l * k + c] * tl[c];
            s = warp_sum(s);
            if (lane == 0) ml[l] = xm[l] + s;
            for (int c = lane; c < k; c += 32) T[l * k + c] *= Dv[c];
          }
          __syncthreads();
          // Xa[l][i] = ml[l] + sum_c T[l][c] M[c][i]   (i = tx + 16 b), written over Xt
          for (int lt2 = ty; lt2 < nlt; lt2 += NTY) {
            double o[TL][TM];
#pragma unroll
            for (int a = 0; a < TL; ++a)
#pragma unroll
              for (int b = 0; b < TM; ++b) o[a][b] = 0.0;
            for (int c = 0; c < k; ++c) {
              double tv[TL], mv[TM];
#pragma unroll
              for (int a = 0; a < TL; ++a) { const int l = lt2 + nlt * a; tv[a] = l < nl ? T[l * k + c] : 0.0; }
#pragma unroll
              for (int b = 0; b < TM; ++b) { const int i = tx + 16 * b; mv[b] = i < k ? M[c * ks + i] : 0.0; }
#pragma unroll
              for (int a = 0; a < TL; ++a)
#pragma unroll
                for (int b = 0; b < TM; ++b) o[a][b] = fma(tv[a], mv[b], o[a][b]);
            }
#pragma unroll
            for (int a = 0; a < TL; ++a)
#pragma unroll
              for (int b = 0; b < TM; ++b) {
                const int l = lt2 + nlt * a, i = tx + 16 * b;
                if (l < nl && i < k) Xt[l * k + i] = ml[l] + o[a][b];
              }
          }
          __syncthreads();
          for (int e = tid; e < nl * k; e += NT) T[e] = Xt[e];
          __syncthreads();
        }
        for (int e = tid; e < nl * k; e += NT) Xg[(long long)l0 * k + e] = T[e];
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
    }  // lt
    if (tid == 0) {
      if (!redo) {
        atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)col_npl);
        atomicMax(&P.stats[1], col_npl);
        atomicAdd((unsigned long long*)&P.stats[5], 1ull);
      }
      else atomicAdd((unsigned long long*)&P.stats[6], 1ull);
      atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)col_sweeps);
      atomicMax(&P.stats[3], (long long)col_sweeps);
    }
  }
}

static size_t v2_smem_bytes(int k, int lch) {
  const size_t ks = (size_t)(k | 1);
  const size_t usz = std::max((size_t)V2_PCH * k + V2_PCH, (size_t)2 * lch * k);
  const size_t dbl = (size_t)k * ks + usz + 3 * (size_t)k + 2 * (size_t)lch + V2_SELCAP;
  return dbl * 8 + (size_t)V2_SELCAP * 4 + 32 * 4 + 4 * 4 + 16;
}
