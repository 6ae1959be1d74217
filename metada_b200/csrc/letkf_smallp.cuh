// letkf_smallp.cuh -- CANONICAL transform in OBSERVATION space for transforms with few local
// observations (p <= SP_PMAX and 2 p <= k), one WARP per transform.
//
// With M = the p x k matrix of weighted local rows sqrt(rho/sigma^2) Y', s = (k-1)/infl and
// S = M M^T = V diag(l) V^T (p x p), every function of A = s I + M^T M is
//     f(A) = f(s) I + M^T V diag((f(s + l_i) - f(s)) / l_i) V^T M,
// so   A^{-1/2} = s^{-1/2} I + M^T Phi M,  Phi = V diag(phi_i) V^T,
//                 phi_i = -1 / (a_i^{1/2} s^{1/2} (a_i^{1/2} + s^{1/2})),  a_i = s + l_i   (no cancellation, finite at l = 0)
//      w = A^{-1} M^T dw = M^T u,          u = V diag(1 / a_i) V^T dw                   (push-through identity)
// and a level's update is  q = M x'^T,  xa_j = xbar + q.u + sqrt(k-1) (s^{-1/2} x'_j + (Phi q)^T M[:, j]):
// O(p^2 k + p^3) for the transform and O(p k) per level instead of ~20 k^3-products -- three orders
// of magnitude less arithmetic for the C4 shape with per-level localisation (p ~ 9, k = 128), where
// the k x k route spends 280 us of a whole SM per transform.  The p x p eigenproblem is a two-sided
// Jacobi with parallel (round-robin) ordering in the warp's shared-memory tile; M is never staged: its rows (1 KB, contiguous)
// are re-read through L1/L2.  Same transform as the k-space kernels to rounding (the symmetric
// square root is unique); the packed Newton-Schulz kernel hands over the transforms that qualify
// (ColParams.small_items), this kernel consumes the list.
#pragma once
#include "letkf_kernels.cuh"

#define SP_PMAX 24    // first choice: 11 KB of shared memory per warp
#define SP_PMAX2 32   // second pass over the transforms with 24 < p <= 32 (one lane per row still): 19 KB per warp
#define SP_WARPS 8
// warps per CTA of the per-level passes: as many resident warps as shared memory allows (the code is latency bound,
// a serial Jacobi per warp) -- 2 CTAs x 10 warps per SM at SP_PMAX (round 1: 2 x 8), 1 x 11 at SP_PMAX2
#define SP_WARPS_P1 10
#define SP_WARPS_P2 11
#define SP_MAXR 4     // k <= 128: members per lane

template <int PMAX>
struct SpWarpSmemT {
  double S[PMAX][PMAX + 1];
  double V[PMAX][PMAX + 1];
  double wgt[PMAX], dw[PMAX], u[PMAX], phi[PMAX], q[PMAX], r[PMAX];
  double2 pcs[PMAX / 2 + 2];             // rotations of one Jacobi round: (c, s) ...
  int pab[PMAX / 2 + 2];                 // ... and the index pair a | b << 8 (one identity entry pads an odd count)
  int row[PMAX];
  int pad[8];
};
using SpWarpSmem = SpWarpSmemT<SP_PMAX>;

// One transform (col, lt) by the calling warp.  Returns 0: done in observation space (sweeps_out =
// Jacobi sweeps), 1: no observation in reach, perturbations inflated, 2: too many local observations
// for this route (nothing written).  p_out = number of local observations.
template <bool EXT, int PMAX = SP_PMAX>
__device__ __forceinline__ int sp_transform(const ColParams& P, SpWarpSmemT<PMAX>& W, long long col, int lt, int lane,
                                            int& p_out, int& sweeps_out) {
  const int k = P.k, nz = P.nz, nr = (k + 31) >> 5;
  const bool per_level = P.radius_v > 0.0;
  const int R = index_reach<EXT>(P.iv, P.radius);
  const double km1 = (double)(k - 1), s = km1 / P.inflation, sW = sqrt(km1), rs = 1.0 / sqrt(s);
  const int lx = (int)(col % P.nx), ly = (int)(col / P.nx);
  int gx = P.gx0 + lx, gy = P.gy0 + ly;
  index_col_coords<EXT>(P.iv, col, gx, gy);
  double* Xg = P.X + col * nz * k;

    // ---- selection (same predicate and weights as the other column kernels)
  int p = 0;
  bool overflow = false;
  int cy0 = 0, cy1 = -1;
  index_cy_range(P.iv, gy, R, cy0, cy1);
  for (int cy = cy0; cy <= cy1; ++cy) {
    int rb, re;
    index_row_range(P.iv, gx, R, cy, rb, re);
    for (int a0 = rb; a0 < re; a0 += 32) {
      const int a = a0 + lane;
      bool sel = false;
      double sq = 0.0, sd = 0.0;
      int orow = 0;
      if (a < re) {
        double dist;
        sel = index_within<EXT>(P.iv, col, a, gx, gy, P.radius, &dist);
        double dv = 0.0;
        if (sel && per_level) {
          dv = fabs((double)(P.iv.sz[a] - index_level<EXT>(P.iv, lt)));
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
          sd = sq > 0.0 ? sq * P.d[orow] : 0.0;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, sel);
      const int pos = p + __popc(bal & ((1u << lane) - 1u));
      if (sel && pos < PMAX) { W.row[pos] = orow; W.wgt[pos] = sq; W.dw[pos] = sd; }
      p += __popc(bal);
      if (p > PMAX) overflow = true;
    }
  }
  __syncwarp();
  p_out = p;
  if (p == 0) {                 // no observation in reach: inflate the perturbations (LETKF.hpp:167-190)
    const double f = sqrt(P.inflation);
    const int lb = per_level ? lt : 0, le = per_level ? lt + 1 : nz;
    for (int l = lb; l < le; ++l) {
      double* x = Xg + (long long)l * k;
      double xv[SP_MAXR], sum = 0.0;
#pragma unroll
      for (int r = 0; r < SP_MAXR; ++r) { xv[r] = (r < nr && lane + 32 * r < k) ? x[lane + 32 * r] : 0.0; sum += xv[r]; }
      const double xbar = warp_sum(sum) / (double)k;
      double msum = 0.0;
#pragma unroll
      for (int r = 0; r < SP_MAXR; ++r)
        if (r < nr && lane + 32 * r < k) { const double v = xbar + (xv[r] - xbar) * f; x[lane + 32 * r] = v; msum += v; }
      if (P.mean_out) { msum = warp_sum(msum); if (lane == 0) P.mean_out[col * nz + l] = msum * (1.0 / (double)k); }
    }
    return 1;
  }
  if (overflow || 2 * p > k) return 2;    // not for this kernel

  // ---- S = M M^T on the FP64 tensor path.  Lane (g, t) of a k-step reads element [8 I + g][4 kk + t] of the NRT row
  // tiles straight from global memory (32 contiguous bytes per row: whole sectors): that one register is the A
  // fragment of tile row I and, S being M times ITS OWN transpose, the B fragment of tile column I as well, so a
  // k-step is NRT independent loads and NRT (NRT + 1) / 2 MMAs.  (Round 1 formed the p (p + 1) / 2 entries one at
  // a time -- four loads from L2, a warp reduction and a store each, ~1 k cycles per entry in a serial chain: for
  // p = 18 as long as the whole Jacobi.)  Row weights are applied to the finished entries.
  {
    constexpr int NRT = PMAX / 8;
    const int g = lane >> 2, t = lane & 3;
    double sacc[NRT][NRT][2];
    const double* rowp[NRT];
    bool rowok[NRT];
#pragma unroll
    for (int I = 0; I < NRT; ++I) {
      rowok[I] = 8 * I + g < p;
      rowp[I] = P.Yp + (long long)W.row[rowok[I] ? 8 * I + g : 0] * k + t;
#pragma unroll
      for (int J = 0; J < NRT; ++J) { sacc[I][J][0] = 0.0; sacc[I][J][1] = 0.0; }
    }
    const int nrt = (p + 7) >> 3;
#pragma unroll 4
    for (int kk = 0; kk < k; kk += 4) {
      double f[NRT];
#pragma unroll
      for (int I = 0; I < NRT; ++I) f[I] = (I < nrt && rowok[I] && kk + t < k) ? rowp[I][kk] : 0.0;
#pragma unroll
      for (int I = 0; I < NRT; ++I)
#pragma unroll
        for (int J = I; J < NRT; ++J)
          if (J < nrt)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(sacc[I][J][0]), "+d"(sacc[I][J][1]) : "d"(f[I]), "d"(f[J]));
    }
#pragma unroll
    for (int I = 0; I < NRT; ++I)
#pragma unroll
      for (int J = I; J < NRT; ++J) {
        const int a = 8 * I + g, b0 = 8 * J + 2 * t;
        if (J < nrt && a < p) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (b0 + e < p) {
              const double v = sacc[I][J][e] * W.wgt[a] * W.wgt[b0 + e];
              W.S[a][b0 + e] = v;
              if (I != J) W.S[b0 + e][a] = v;          // (a diagonal tile holds both halves itself)
            }
        }
      }
  }
  if (lane < p) for (int i = 0; i < p; ++i) W.V[i][lane] = (i == lane) ? 1.0 : 0.0;
  __syncwarp();

  // ---- two-sided Jacobi on S (p x p), eigenvectors in the columns of V.  Parallel ordering: a round
  // rotates p/2 DISJOINT index pairs at once (round-robin tournament, p - 1 rounds per sweep); lane l
  // computes the rotation of pair l, then every lane applies all of them to its row (S J, V J) and to
  // its column (J^T S).  The scalar sqrt/div chain of a rotation is the latency that matters here, and
  // this way it is paid p - 1 times per sweep instead of p (p - 1) / 2.
  const int n = (p + 1) & ~1, half = n >> 1;
  if (lane == half) { W.pab[lane] = 0; W.pcs[lane] = make_double2(1.0, 0.0); }   // identity entry that pads an odd pair count
  int sweeps = 0;
  for (; sweeps < 30; ++sweeps) {
    double off = 0.0, dia = 0.0;
    if (lane < p)                                             // (column `lane` of every row: no index arithmetic)
      for (int i = 0; i < p; ++i) {
        const double v = W.S[i][lane];
        if (i == lane) dia = fma(v, v, dia); else off = fma(v, v, off);
      }
    off = warp_sum(off); dia = warp_sum(dia);
    if (!(off > 1e-26 * dia) || off == 0.0) break;      // off-diagonal mass below 1e-13 of the diagonal
    for (int r = 0; r < n - 1; ++r) {
      if (lane < half) {
        int a = r + lane, b = r - lane;                          // (mod n - 1 without divisions: r < n - 1, lane < n / 2)
        if (a >= n - 1) a -= n - 1;
        if (b < 0) b += n - 1;
        if (lane == 0) a = n - 1;
        if (a > b) { const int t_ = a; a = b; b = t_; }
        double c = 1.0, sn = 0.0;
        if (b < p) {                                           // (index p is the padding of an odd p)
          const double apq = W.S[a][b], app = W.S[a][a], aqq = W.S[b][b];
          // (a pair whose off-diagonal entry is below 1e-16 of its diagonal neighbourhood is rotated by less than a
          // unit round-off: skipped, and the update loops skip it too -- most pairs of the last sweep; apq^2 must not underflow)
          if (fabs(apq) > 1e-140 && apq * apq > 1e-32 * fabs(app * aqq)) {
            // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = al / apq, al = (S_bb - S_aa) / 2, with numerator
            // and denominator multiplied by |apq|: one reciprocal square root, one division, one reciprocal square
            // root in the dependent chain (the textbook form has three divisions and two square roots -- this chain,
            // run by p / 2 lanes, was 30 % of the per-level first pass)
            const double al = 0.5 * (aqq - app);
            const double h2 = fma(al, al, apq * apq);
            const double hyp = h2 * rsqrt(h2);
            double t = fabs(apq) / (fabs(al) + hyp);
            if (al != 0.0 && ((al < 0.0) != (apq < 0.0))) t = -t;
            c = rsqrt(fma(t, t, 1.0)); sn = t * c;
          }
        }
        W.pab[lane] = a | (b << 8); W.pcs[lane] = make_double2(c, sn);
      }
      __syncwarp();
      // S <- S J, V <- V J (row `lane`), then S <- J^T S (column `lane`).  The pairs of a round are disjoint, so two
      // of them are loaded before either is stored: the compiler cannot know that and would chain every pair's
      // loads behind the previous pair's stores.
      if (lane < p) {
        for (int l = 0; l < half; l += 2) {
          const double2 cs0 = W.pcs[l], cs1 = W.pcs[l + 1];
          const int ab0 = W.pab[l], ab1 = W.pab[l + 1];
          const double c0 = cs0.x, sn0 = cs0.y, c1 = cs1.x, sn1 = cs1.y;
          const int a0 = ab0 & 255, b0 = ab0 >> 8, a1 = ab1 & 255, b1 = ab1 >> 8;
          const double sa0 = W.S[lane][a0], sb0 = W.S[lane][b0], va0 = W.V[lane][a0], vb0 = W.V[lane][b0];
          const double sa1 = W.S[lane][a1], sb1 = W.S[lane][b1], va1 = W.V[lane][a1], vb1 = W.V[lane][b1];
          if (sn0 != 0.0) {
            W.S[lane][a0] = c0 * sa0 - sn0 * sb0; W.S[lane][b0] = sn0 * sa0 + c0 * sb0;
            W.V[lane][a0] = c0 * va0 - sn0 * vb0; W.V[lane][b0] = sn0 * va0 + c0 * vb0;
          }
          if (sn1 != 0.0) {
            W.S[lane][a1] = c1 * sa1 - sn1 * sb1; W.S[lane][b1] = sn1 * sa1 + c1 * sb1;
            W.V[lane][a1] = c1 * va1 - sn1 * vb1; W.V[lane][b1] = sn1 * va1 + c1 * vb1;
          }
        }
      }
      __syncwarp();
      if (lane < p) {
        for (int l = 0; l < half; l += 2) {
          const double2 cs0 = W.pcs[l], cs1 = W.pcs[l + 1];
          const int ab0 = W.pab[l], ab1 = W.pab[l + 1];
          const double c0 = cs0.x, sn0 = cs0.y, c1 = cs1.x, sn1 = cs1.y;
          const int a0 = ab0 & 255, b0 = ab0 >> 8, a1 = ab1 & 255, b1 = ab1 >> 8;
          const double sa0 = W.S[a0][lane], sb0 = W.S[b0][lane], sa1 = W.S[a1][lane], sb1 = W.S[b1][lane];
          if (sn0 != 0.0) { W.S[a0][lane] = c0 * sa0 - sn0 * sb0; W.S[b0][lane] = sn0 * sa0 + c0 * sb0; }
          if (sn1 != 0.0) { W.S[a1][lane] = c1 * sa1 - sn1 * sb1; W.S[b1][lane] = sn1 * sa1 + c1 * sb1; }
        }
      }
      __syncwarp();
    }
  }

  // ---- phi_i, u = V diag(1/a) V^T dw
  if (lane < p) {
    const double l = fmax(W.S[lane][lane], 0.0), a = s + l, ra = sqrt(a), r0 = sqrt(s);
    W.phi[lane] = -1.0 / (ra * r0 * (ra + r0));
    double tv = 0.0;
    for (int b = 0; b < p; ++b) tv = fma(W.V[b][lane], W.dw[b], tv);   // (V^T dw)_lane
    W.q[lane] = tv / a;
  }
  __syncwarp();
  if (lane < p) {
    double uv = 0.0;
    for (int b = 0; b < p; ++b) uv = fma(W.V[lane][b], W.q[b], uv);
    W.u[lane] = uv;
  }
  __syncwarp();

  // ---- levels
  const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
  for (int l = lev_b; l < lev_e; ++l) {
    double* x = Xg + (long long)l * k;
    double xv[SP_MAXR];
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < SP_MAXR; ++r) {
      xv[r] = (r < nr && lane + 32 * r < k) ? x[lane + 32 * r] : 0.0;
      sum += xv[r];
    }
    const double xbar = warp_sum(sum) / (double)k;
#pragma unroll
    for (int r = 0; r < SP_MAXR; ++r) xv[r] = (r < nr && lane + 32 * r < k) ? xv[r] - xbar : 0.0;
    // q = M x'^T
    for (int a = 0; a < p; ++a) {
      const double* ya = P.Yp + (long long)W.row[a] * k;
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < SP_MAXR; ++r)
        if (r < nr && lane + 32 * r < k) acc = fma(ya[lane + 32 * r], xv[r], acc);
      acc = warp_sum(acc);
      if (lane == 0) W.q[a] = acc * W.wgt[a];
    }
    __syncwarp();
    // c = q . u ;  r = Phi q = V diag(phi) V^T q, scaled by the row weights for the last product
    double cpart = (lane < p) ? W.q[lane] * W.u[lane] : 0.0;
    const double cq = warp_sum(cpart);
    double tq = 0.0;
    if (lane < p) {
      for (int b = 0; b < p; ++b) tq = fma(W.V[b][lane], W.q[b], tq);
      tq *= W.phi[lane];
    }
    __syncwarp();
    if (lane < p) W.S[0][lane] = tq;            // S is dead: reuse its first row as scratch
    __syncwarp();
    if (lane < p) {
      double rv = 0.0;
      for (int b = 0; b < p; ++b) rv = fma(W.V[lane][b], W.S[0][b], rv);
      W.r[lane] = rv * W.wgt[lane];
    }
    __syncwarp();
    double out[SP_MAXR], msum = 0.0;
#pragma unroll
    for (int r = 0; r < SP_MAXR; ++r) out[r] = rs * xv[r];
    for (int a = 0; a < p; ++a) {
      const double* ya = P.Yp + (long long)W.row[a] * k;
      const double ra = W.r[a];
#pragma unroll
      for (int r = 0; r < SP_MAXR; ++r)
        if (r < nr && lane + 32 * r < k) out[r] = fma(ra, ya[lane + 32 * r], out[r]);
    }
#pragma unroll
    for (int r = 0; r < SP_MAXR; ++r)
      if (r < nr && lane + 32 * r < k) {
        const double v = xbar + cq + sW * out[r];
        x[lane + 32 * r] = v;
        msum += v;
      }
    if (P.mean_out) {
      msum = warp_sum(msum);
      if (lane == 0) P.mean_out[col * nz + l] = msum * (1.0 / (double)k);
    }
    __syncwarp();
  }
  sweeps_out = sweeps;
  return 0;
}

// Consumer of the packed kernel's list of small transforms (non-per-level analyses).
template <bool EXT, int PMAX = SP_PMAX, int WARPS = SP_WARPS>
__global__ void __launch_bounds__(WARPS * 32) letkf_smallp_kernel(ColParams P) {
  extern __shared__ __align__(16) unsigned char sp_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SpWarpSmemT<PMAX>& W = reinterpret_cast<SpWarpSmemT<PMAX>*>(sp_smem_raw)[warp];
  const int nxf = P.radius_v > 0.0 ? P.nz : 1;
  const long long nitems = (long long)*P.small_count;
  for (long long it = (long long)blockIdx.x * WARPS + warp; it < nitems; it += (long long)gridDim.x * WARPS) {
    const long long item = P.small_items[it], col = item / nxf;
    int p = 0, sweeps = 0;
    const int rc = sp_transform<EXT, PMAX>(P, W, col, (int)(item - col * nxf), lane, p, sweeps);
    if (lane == 0) {
      if (rc == 2) atomicAdd((unsigned long long*)&P.stats[4], 1ull);   // cannot happen: the producer counted p
      atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)sweeps);
      atomicMax(&P.stats[3], (long long)sweeps);
      if (rc == 0) atomicAdd((unsigned long long*)&P.stats[7], 1ull);
    }
  }
}

// Per-level analyses: FIRST pass over every (column, level) transform, one warp per transform.  Transforms
// with no or few local observations are finished here; those with SP_PMAX < p <= SP_PMAX2 (and 2 p <= k) go to a
// second list, which letkf_smallp_kernel<false, SP_PMAX2, ..> finishes in observation space too (P.small_items);
// the others go to the work list of the packed k-space kernel, which then never spends a 512-thread selection
// on a transform it will not do.  (BASELINE C4, r_v = 5: mean p ~ 18 at interior levels; 94 % of the transforms
// have p <= 24, 99.9 % p <= 32, and a k-space transform at k = 128 costs ~130 us of a whole SM.)
template <int WARPS, bool EXT = false>
__global__ void __launch_bounds__(WARPS * 32) letkf_smallp_classify_kernel(ColParams P) {
  extern __shared__ __align__(16) unsigned char sp_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SpWarpSmem& W = reinterpret_cast<SpWarpSmem*>(sp_smem_raw)[warp];
  const int nxf = P.radius_v > 0.0 ? P.nz : 1;
  const long long ncols = P.cols ? P.ncols : (long long)P.own_nx * P.own_ny;
  // one warp per TRANSFORM (consecutive warps take the levels of one column: their index reads share L1)
  for (long long ti = (long long)blockIdx.x * WARPS + warp; ti < ncols * nxf; ti += (long long)gridDim.x * WARPS) {
    const long long ci = ti / nxf;
    const int lt = (int)(ti - ci * nxf);
    long long col;
    if (P.cols) col = P.cols[ci];
    else col = (ci / P.own_nx) * P.nx + ci % P.own_nx;
    int p = 0, sweeps = 0;
    const int rc = sp_transform<EXT>(P, W, col, lt, lane, p, sweeps);
    if (lane == 0) {
      if (rc == 2) {
        if (P.small_items && p <= SP_PMAX2 && 2 * p <= P.k) {
          const unsigned slot = atomicAdd(P.small_count, 1u);
          P.small_items[slot] = col * nxf + lt;
        } else {
          const unsigned slot = atomicAdd(P.work_count, 1u);
          P.work_items[slot] = col * nxf + lt;
        }
      }
      if (rc == 0) {
        atomicAdd((unsigned long long*)&P.stats[7], 1ull);
        atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)sweeps);
        atomicMax(&P.stats[3], (long long)sweeps);
      }
      if (lt == 0) {
        atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)p);
        atomicMax(&P.stats[1], (long long)p);
        atomicAdd((unsigned long long*)&P.stats[5], 1ull);
      }
    }
    __syncwarp();
  }
}

static size_t smallp_smem_bytes(int pmax = SP_PMAX, int warps = SP_WARPS) {
  return (pmax == SP_PMAX2 ? sizeof(SpWarpSmemT<SP_PMAX2>) : sizeof(SpWarpSmem)) * (size_t)warps;
}
