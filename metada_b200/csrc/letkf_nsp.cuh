// letkf_nsp.cuh -- CANONICAL column kernel, Newton-Schulz square root on PACKED symmetric tiles.
//
// Same mathematics as letkf_ns.cuh (coupled Newton-Schulz from a Chebyshev start, every product on
// the FP64 tensor path), different storage: every iterate is a symmetric polynomial in A, so only
// the nt (nt + 1) / 2 upper-triangular 8 x 8 tiles are kept, each tile a dense 512-byte block.
//   * k = 80: 55 tiles = 28 KB per matrix instead of 54 KB, and the iteration needs three matrices
//     (Z, Y, T; each product is held in the accumulator registers across a barrier and written
//     back in place) instead of four  ->  ~96 KB per CTA, TWO CTAs per SM: one column's
//     selection / gather / update phases and barrier bubbles hide under the other's products.
//   * k = 128: 136 tiles = 70 KB per matrix, three of them fit one SM (the padded square layout
//     does not), so C4-sized ensembles get the tensor path too (16 warps, one CTA per SM).
// A fragment of the logical matrix S[8 I + g][8 K + 4 h + t] is read from tile (I, K) when I <= K
// and, by symmetry, transposed from tile (K, I) otherwise; the column index inside a tile is XORed
// with 4 on rows 2, 3, 6, 7, which makes the straight read (8 rows x 4 doubles), the transposed read
// (4 rows x 8 doubles) and the 16-byte accumulator stores all bank-conflict free without padding.
// Off-diagonal tiles are stored once (no mirrored store).
//
// Forcing symmetry is only stable while cond(A) is moderate (letkf_ns.cuh); a column (or level, with
// per-level transforms) whose rigorous condition bound exceeds NSP_KAPPA_MAX is appended to a
// redo list, which a second launch of the full-product kernel (k <= 80) or the Jacobi kernel
// (k > 80) consumes.
#pragma once
#include <type_traits>
#include <utility>

#include "letkf_ns.cuh"
#include "letkf_smallp.cuh"
#include "ns_schedule_table.h"

// -DNSP_PROFILE: thread 0 of every CTA adds the clock64() ticks it spends per phase to stats[8..15]
// (8 selection, 9 gather + SYRK, 10 norm / A store / start look-up, 11 products, 12 iteration epilogues,
// 13 w = Z Z g, 14 update, 15 whole column); read with mdc_ctx_last_stats.  Development only.
#ifdef NSP_PROFILE
#define NSP_T0() long long nsp_t_ = clock64()
#define NSP_TICK(slot) do { const long long n_ = clock64(); if (threadIdx.x == 0) atomicAdd((unsigned long long*)&nsp_prof_[slot], (unsigned long long)(n_ - nsp_t_)); nsp_t_ = n_; } while (0)
__device__ long long* nsp_prof_;
#else
#define NSP_T0() do {} while (0)
#define NSP_TICK(slot) do {} while (0)
#endif

__host__ __device__ inline int nsp_ntiles(int k) { const int nt = (k + 7) >> 3; return nt * (nt + 1) / 2; }
__device__ __forceinline__ int nsp_row_start(int I, int nt) { return (I * (2 * nt - I + 1)) >> 1; }

// offset (in doubles) of S[i][j] inside a packed matrix; reads the upper copy
__device__ __forceinline__ int nsp_elem(int i, int j, int nt) {
  if (i > j) { const int s = i; i = j; j = s; }
  const int I = i >> 3, J = j >> 3, r = i & 7, c = j & 7;
  return ((nsp_row_start(I, nt) + J - I) << 6) + r * 8 + (c ^ ((r & 2) << 1));
}

struct NspLane { unsigned offd, offt, offc, dh; };   // byte offsets inside a tile (k-half h = 0)
__device__ __forceinline__ NspLane nsp_lane(int lane) {
  const int g = lane >> 2, t = lane & 3;
  NspLane L;
  L.offd = (unsigned)((g * 8 + (t ^ ((g & 2) << 1))) * 8);          // element (g, t); h = 1: ^ 32 = + dh
  L.dh = (g & 2) ? 0xffffffe0u : 32u;
  L.offt = (unsigned)((t * 8 + (g ^ ((t & 2) << 1))) * 8);          // element (t, g); h = 1: + 256
  L.offc = (unsigned)((g * 8 + ((2 * t) ^ ((g & 2) << 1))) * 8);    // accumulator pair (g, 2t), (g, 2t+1)
  return L;
}
// Fragment element S[8 I + g][8 K + 4 h + t] lives at  tile(I, K) * 512 + lane part, with
//   K <  I: tile (K, I) read transposed, lane part offt + 256 h, next tile (K -> K + 1) + (nt - K - 1)
//   K >= I: tile (I, K) read straight,   lane part offd + dh h,  next tile + 1
// and tile(I, 0) = I in both cases, so a running tile offset needs one select + add per K step.
struct NspWalk {
  unsigned tile;   // byte offset of the current tile
  int I;
  __device__ __forceinline__ void start(int I_) { I = I_; tile = (unsigned)I_ << 9; }
  __device__ __forceinline__ unsigned addr(unsigned base, int K, int h, const NspLane& L) const {
    const bool tr = K < I;
    return base + tile + (tr ? L.offt : L.offd) + (h ? (tr ? 256u : L.dh) : 0u);
  }
  __device__ __forceinline__ void next(int K, int nt) { tile += (K < I) ? ((unsigned)(nt - K - 1) << 9) : 512u; }
};

template <int NTH>
__device__ __forceinline__ double nsp_block_reduce(double v, bool is_max, double* red /*[NTH/32]*/) {
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
#pragma unroll
  for (int w = 1; w < NTH / 32; ++w) r = is_max ? fmax(r, red[w]) : r + red[w];
  return r;
}

// upper-triangular tiles of this warp: contiguous chunk of the row-major enumeration, plus the byte
// offset of each tile's accumulator pair in packed storage
template <int NTW>
struct NspTiles {
  int ti[NTW], tj[NTW];
  int n;
};
__device__ __forceinline__ unsigned nsp_cbase(int ti, int tj, int nt, const NspLane& L) {
  return (((unsigned)(nsp_row_start(ti, nt) + tj - ti)) << 9) + L.offc;
}
template <int NTW, int NTH>
__device__ __forceinline__ NspTiles<NTW> nsp_tiles(int nt, int warp) {
  const int E = nt * (nt + 1) / 2;
  const int e0 = (E * warp) / (NTH / 32), e1 = (E * (warp + 1)) / (NTH / 32);
  NspTiles<NTW> w;
  w.n = e1 - e0;
  int i = 0, rowstart = 0;
  while (e0 >= rowstart + (nt - i)) { rowstart += nt - i; ++i; }
  int j = i + (e0 - rowstart);
#pragma unroll
  for (int n = 0; n < NTW; ++n) {
    w.ti[n] = i; w.tj[n] = j;
    if (n + 1 < w.n) { if (++j == nt) { ++i; j = i; } }
  }
  return w;
}

__device__ __forceinline__ void sts_f64x2(unsigned addr, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
}

#define NSP_DMMA(acc, a, b)                                                                      \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" \
               : "+d"((acc)[0]), "+d"((acc)[1])                                                 \
               : "d"(a), "d"(b))

// ---- The Newton-Schulz products proper: ONE code site per kernel (the iteration is a small state
// machine around it), specialised per warp at compile time.  The tile count NT and the warp index W
// fix the warp's tile list, hence every fragment address: each load is  base register + immediate
// (six base registers: straight h = 0 / h = 1 and transposed, for P and for Q), no address
// arithmetic, no predicates, A fragments shared between the tiles of one tile row at compile time.
// (Double-buffering the fragments in registers spills under the 128-register cap of two CTAs per SM.)
// (A rolled generic version with a running tile walk per operand -- NspWalk, still used by the update
// product -- measured ~17 instructions per DMMA and no faster than the padded one-CTA kernel.)
template <int B, int E, typename F>
__device__ __forceinline__ void nsp_static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    nsp_static_for<B + 1, E>(f);
  }
}
__host__ __device__ constexpr int nsp_tile_i(int nt, int e) { int i = 0, rs = 0; while (e >= rs + (nt - i)) { rs += nt - i; ++i; } return i; }
__host__ __device__ constexpr int nsp_tile_j(int nt, int e) { int i = 0, rs = 0; while (e >= rs + (nt - i)) { rs += nt - i; ++i; } return i + (e - rs); }
__host__ __device__ constexpr int nsp_rs(int I, int nt) { return (I * (2 * nt - I + 1)) / 2; }

// The operands are read with plain loads through pointers derived from the dynamic shared array
// (provably shared, so they compile to LDS with immediate offsets): with inline-asm loads ptxas
// funnels every B fragment through one register and serialises load -> MMA -> load.
__device__ __forceinline__ unsigned char* nsp_smem() {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  return smem_raw;
}
// fragment element S[8 I + g][8 K + 4 h + t]: d = matrix + straight lane offset of this k-half,
// tr = matrix + transposed lane offset of this k-half
template <int NT, int I, int K>
__device__ __forceinline__ double nsp_frag_fixed(const double* d, const double* tr) {
  if constexpr (I <= K) return d[(nsp_rs(I, NT) + K - I) * 64];
  else return tr[(nsp_rs(K, NT) + I - K) * 64];
}

// The k-half loop (h) stays rolled: the body is then half as long, and the eight warp variants of
// NT = 10 together (20 KB) stay inside the 32 KB L1.5 instruction cache; fully unrolled (41 KB) the
// same code measured no faster than the generic walk.
template <int NT, int NW, int NTW, int W>
__device__ __forceinline__ void nsp_mm_warp(unsigned pbase, unsigned qbase, const NspLane& L, double (&acc)[NTW][2]) {
  constexpr int E = NT * (NT + 1) / 2, e0 = (E * W) / NW, e1 = (E * (W + 1)) / NW, N = e1 - e0;
  const unsigned char* sm = nsp_smem();
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const double* pd = reinterpret_cast<const double*>(sm + pbase + L.offd + (h ? L.dh : 0u));
    const double* ptr = reinterpret_cast<const double*>(sm + pbase + L.offt + (h << 8));
    const double* qd = reinterpret_cast<const double*>(sm + qbase + L.offd + (h ? L.dh : 0u));
    const double* qtr = reinterpret_cast<const double*>(sm + qbase + L.offt + (h << 8));
    nsp_static_for<0, NT>([&](auto K_c) {
      constexpr int K = decltype(K_c)::value;
      double a[NTW], b[NTW];
      nsp_static_for<0, N>([&](auto n_c) {
        constexpr int n = decltype(n_c)::value;
        constexpr int ti = nsp_tile_i(NT, e0 + n), tj = nsp_tile_j(NT, e0 + n);
        if constexpr (n > 0 && ti == nsp_tile_i(NT, e0 + (n > 0 ? n - 1 : 0))) a[n] = a[n - 1];
        else a[n] = nsp_frag_fixed<NT, ti, K>(pd, ptr);
        b[n] = nsp_frag_fixed<NT, tj, K>(qd, qtr);
      });
      nsp_static_for<0, N>([&](auto n_c) {
        constexpr int n = decltype(n_c)::value;
        NSP_DMMA(acc[n], a[n], b[n]);
      });
    });
  }
}

template <int NT, int NW, int NTW, int... Ws>
__device__ __forceinline__ void nsp_mm_dispatch(std::integer_sequence<int, Ws...>, int warp, unsigned pbase,
                                                unsigned qbase, const NspLane& L, double (&acc)[NTW][2]) {
  ((warp == Ws ? (nsp_mm_warp<NT, NW, NTW, Ws>(pbase, qbase, L, acc), 0) : 0), ...);
}

// Generic product for the larger tile counts (their specialised bodies would not fit the
// instruction cache): a running tile walk per operand, ~17 instructions per DMMA.  Warps with one
// tile fewer than NTW repeat their last tile (never stored) so that no MMA is predicated.
template <int NTW>
__device__ __forceinline__ void nsp_mm_walk(unsigned pbase, unsigned qbase, int nt, const NspTiles<NTW>& w,
                                            const NspLane& L, double (&acc)[NTW][2]) {
  NspWalk pw[NTW], qw[NTW];
#pragma unroll
  for (int n = 0; n < NTW; ++n) { pw[n].start(w.ti[n]); qw[n].start(w.tj[n]); }
#pragma unroll 1
  for (int K = 0; K < nt; ++K) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double a[NTW], b[NTW];
#pragma unroll
      for (int n = 0; n < NTW; ++n) {
        if (n == 0 || w.ti[n] != w.ti[n - 1]) a[n] = lds_f64(pw[n].addr(pbase, K, h, L));
        else a[n] = a[n - 1];
        b[n] = lds_f64(qw[n].addr(qbase, K, h, L));
      }
#pragma unroll
      for (int n = 0; n < NTW; ++n) NSP_DMMA(acc[n], a[n], b[n]);
    }
#pragma unroll
    for (int n = 0; n < NTW; ++n) { pw[n].next(K, nt); qw[n].next(K, nt); }
  }
}

#define NSP_FIXED_MAX_NT 10
#define NSP_LCH_MAX 32   /* levels per update chunk */
template <int NT, int NW, int NTW>
__device__ __forceinline__ void nsp_mm_any(unsigned pbase, unsigned qbase, int warp, const NspTiles<NTW>& w,
                                           const NspLane& L, double (&acc)[NTW][2]) {
#pragma unroll
  for (int n = 0; n < NTW; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
  if constexpr (NT <= NSP_FIXED_MAX_NT) {
    // byte offsets from the start of dynamic shared memory
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(nsp_smem());
    nsp_mm_dispatch<NT, NW, NTW>(std::make_integer_sequence<int, NW>{}, warp, pbase - s0, qbase - s0, L, acc);
  } else {
    nsp_mm_walk<NTW>(pbase, qbase, NT, w, L, acc);
  }
}

template <int NTW>
__device__ __forceinline__ void nsp_store(unsigned dbase, int nt, const NspTiles<NTW>& w, const NspLane& L,
                                          double (&acc)[NTW][2]) {
#pragma unroll
  for (int n = 0; n < NTW; ++n)
    if (n < w.n) sts_f64x2(dbase + nsp_cbase(w.ti[n], w.tj[n], nt, L), acc[n][0], acc[n][1]);
}

template <int NTA, int NTH>
__device__ __forceinline__ NsTiles<NTA> nsp_rect_tiles(int ntr, int nt, int warp) {
  const int E = ntr * nt;
  const int e0 = (E * warp) / (NTH / 32), e1 = (E * (warp + 1)) / (NTH / 32);
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

// ---- schedule look-ups (ns_schedule_table.h): all threads compute the same index (constant-memory broadcasts)
__device__ __forceinline__ int nss_start_index(double kappa) {
  // kappa grid: kappa_i - 1 = 1e-3 * 1.12^i; smallest i with kappa_i >= kappa
  int i = (int)ceilf(__log2f(fmaxf((float)(kappa - 1.0), 1e-3f) * 1e3f) * (1.0f / 0.16349873f));
  i = max(0, min(i, NSS_NKAPPA - 1));
  while (i > 0 && nss_starts[i - 1].kappa >= kappa) --i;
  while (i < NSS_NKAPPA - 1 && nss_starts[i].kappa < kappa) ++i;
  return (nss_starts[i].kappa >= kappa) ? i : -1;
}
__device__ __forceinline__ int nss_step_index(double rho) {
  // rho grid (descending): 2 rho_i / (1 - rho_i) = 4000 * 0.88^i; largest i with rho_i >= rho
  if (!(rho <= nss_steps[0].rho)) return -1;
  const float gq = (float)(2.0 * rho / (1.0 - rho));
  int i = (int)floorf(__log2f(fmaxf(gq, 1e-12f) * (1.0f / 4000.0f)) * (1.0f / -0.18442457f));
  i = max(0, min(i, NSS_NRHO - 1));
  while (i < NSS_NRHO - 1 && nss_steps[i + 1].rho >= rho) ++i;
  while (i > 0 && nss_steps[i].rho < rho) --i;
  return i;
}

// largest condition bound (Schatten-4 bound of the spectrum / shift) the packed kernel takes; beyond it the
// transform goes to the redo list.  Symmetric-tile products assume the iterates commute; with the short composite
// minimax schedule (13 - 20 products) the rounding defect stays small much longer than with plain Newton-Schulz:
// error of Z against the eigen-decomposition (numpy emulation of these very tile products on C5-like matrices)
// 5e-15 at cond 50, 7e-15 at 200, 1.5e-14 at 530, 3.3e-14 at 1200, 1.4e-13 at 3500.
#define NSP_KAPPA_MAX 2000.0

// Z <- A^{-1/2} for the A held in the T buffer, by the composite minimax polynomial iteration of
// tools/gen_ns_schedule.py: state Z and the residual E = I - Z^2 A (in the Y buffer), spectrum(E) in [-rho, rho];
//   start  (degree 0..2 in A, chosen by the condition bound kappa):  Z0 = q(A), E0 = I - Z0 (A Z0)
//   stage  (degree d = 1..3):  T = sum c_i E^i,  Z <- Z T,  E <- I - T^2 + E T^2       d + 2 products
//   finish (degree f = 1..3):  Z <- Z sum c_i E^i                                       f products
// with the minimax coefficients for the interval the spectrum is known to lie in and the degree sequence that
// minimises the product count (13 - 14 products at C5's conditioning; Chebyshev start + Newton-Schulz + series
// finish took 17 - 20).  rho is tracked a priori (rigorous while spectrum(A) is inside [shift, shift kappa]) and
// cross-checked against the measured ||E||_F after every stage.
// Deliberately NOT inlined: the column loop around it keeps ~60 registers of state alive, and under the
// 128-register cap ptxas then serialises every fragment load behind the MMA that frees its register; as a
// separate function the products get the whole budget (the caller's state is saved once per column).
// Returns the number of k x k products used, -1 if the iteration failed, -2 if kappa is beyond NSP_KAPPA_MAX.
template <int NT, int NTH>
__device__ __noinline__ int nsp_inverse_sqrt(double* Zp, double shift, double fro, int k) {
  constexpr int NW = NTH / 32, NTW = (NT * (NT + 1) / 2 + NW - 1) / NW, nt = NT, kp = 8 * NT;
  constexpr int msz = NT * (NT + 1) / 2 * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  double* Yp = Zp + msz;
  double* Tp = Yp + msz;
  double* red = Tp + msz + 3 * kp + 2 * NSP_LCH_MAX;
  const unsigned zs = (unsigned)__cvta_generic_to_shared(Zp), ys = zs + msz * 8, ts = ys + msz * 8;
  const NspLane L = nsp_lane(lane);
  const NspTiles<NTW> st = nsp_tiles<NTW, NTH>(nt, warp);
  // the lane's accumulator pair of tile n sits at byte offset cb(n) of every packed matrix; dg(n): 0 = off the
  // diagonal, 1 = first element on it, 2 = second
  auto cb = [&](int n) { return nsp_cbase(st.ti[n], st.tj[n], nt, L); };
  auto dg = [&](int n) { const int i = st.ti[n] * 8 + g, j = st.tj[n] * 8 + 2 * t; return i == j ? 1 : (i == j + 1 ? 2 : 0); };
  auto own = [&](const double* M, int n) { return *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(M) + cb(n)); };
  // The iteration is a small state machine around the SINGLE product site (nsp_mm_any; its per-warp specialised
  // bodies must exist once for the instruction cache).  Operands by op:
  enum { OP_A2, OP_Y0, OP_M0, OP_E2, OP_E3, OP_ZT, OP_T2, OP_ET };
  double acc[NTW][2];
  double rho_ap = 0.0;
  int op = OP_A2, nprod = 0, step = 0;     // step: index of the current stage / finish in nss_steps
  int rc = -1;
  // next step from the residual bound: loads the coefficients, returns the first op of the step (-1: failure)
  auto plan = [&](double rmeas) -> int {
    if (!(rmeas < 1e6)) return -1;                                  // NaN / divergence guard
    const int j = nss_step_index(fmin(rho_ap, rmeas));
    if (j < 0) return -1;
    step = j;
    rho_ap = nss_steps[j].rho_out * 1.002;
    if (nss_steps[j].kind % 10 == 1) {                              // T = c0 I + c1 E straight into the T buffer
      const double c0 = nss_steps[j].c[0], c1 = nss_steps[j].c[1];
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const double2 e = own(Yp, n);
          const int d = dg(n);
          sts_f64x2(ts + cb(n), fma(c1, e.x, d == 1 ? c0 : 0.0), fma(c1, e.y, d == 2 ? c0 : 0.0));
        }
      __syncthreads();
      return OP_ZT;
    }
    return OP_E2;
  };
  NSP_T0();
#pragma unroll 1
  while (true) {
    const unsigned pb = (op == OP_A2 || op == OP_Y0 || op == OP_T2) ? ts : (op == OP_M0 || op == OP_ZT) ? zs : ys;
    const unsigned qb = (op == OP_Y0) ? zs : (op == OP_M0 || op == OP_E2) ? ys : ts;
    NSP_TICK(12);
    nsp_mm_any<NT, NW, NTW>(pb, qb, warp, st, L, acc);
    NSP_TICK(11);
    ++nprod;
    if (op == OP_A2) {
      // tighter upper end of the spectrum from the product just made: lmax(C)^2 <= ||C^2||_F
      // (Schatten-4 norm of C; C^2 = A^2 - 2 shift A + shift^2 I), typically 2-3x below ||C||_F
      double f4 = 0.0;
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const int i = st.ti[n] * 8 + g, j = st.tj[n] * 8 + 2 * t, d = dg(n);
          const double2 a = own(Tp, n);
          const double w = (st.ti[n] != st.tj[n]) ? 2.0 : 1.0;
          const double q0 = acc[n][0] - 2.0 * shift * a.x + (d == 1 ? shift * shift : 0.0);
          const double q1 = acc[n][1] - 2.0 * shift * a.y + (d == 2 ? shift * shift : 0.0);
          if (i < k && j < k) f4 = fma(w * q0, q0, f4);
          if (i < k && j + 1 < k) f4 = fma(w * q1, q1, f4);
        }
      f4 = nsp_block_reduce<NTH>(f4, false, red);
      const double hi = shift + fro;
      const double kappa = fmax(fmin(hi, shift + sqrt(sqrt(f4) + 1e-13 * hi * hi)) / shift, 1.0) * (1.0 + 1e-9);
      const int si = (kappa <= NSP_KAPPA_MAX) ? nss_start_index(kappa) : -1;
      if (si < 0) { rc = -2; break; }
      const int sdeg = nss_starts[si].degree;
      rho_ap = nss_starts[si].rho0 * 1.002;
      const double rs = rsqrt(shift), is = 1.0 / shift;
      const double z0 = nss_starts[si].a[0] * rs, z1 = nss_starts[si].a[1] * rs * is, z2 = nss_starts[si].a[2] * rs * is * is;
      // Z0 = q(A) -> Z buffer; degree 0: E0 = I - z0^2 A, degree 1: Y0 = A Z0 = z0 A + z1 A^2 -> Y buffer (no product)
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const double2 a = own(Tp, n);
          const int d = dg(n);
          const double dx = (d == 1 ? 1.0 : 0.0), dy = (d == 2 ? 1.0 : 0.0);
          sts_f64x2(zs + cb(n), fma(z2, acc[n][0], fma(z1, a.x, z0 * dx)), fma(z2, acc[n][1], fma(z1, a.y, z0 * dy)));
          if (sdeg == 0) sts_f64x2(ys + cb(n), dx - z0 * z0 * a.x, dy - z0 * z0 * a.y);
          else if (sdeg == 1) sts_f64x2(ys + cb(n), fma(z1, acc[n][0], z0 * a.x), fma(z1, acc[n][1], z0 * a.y));
        }
      __syncthreads();
      if (sdeg == 0) {
        // (||E0||_F is not measured: rho0 = (kappa - 1) / (kappa + 1) is exact for the bound)
        op = plan(rho_ap);
        if (op < 0) break;
      } else {
        op = (sdeg == 1) ? OP_M0 : OP_Y0;
      }
    } else if (op == OP_Y0) {
      nsp_store<NTW>(ys, nt, st, L, acc);                       // Y0 = A Z0 (the Y buffer is free)
      __syncthreads();
      op = OP_M0;
    } else if (op == OP_M0 || op == OP_ET) {
      // E = I - Z0 Y0, or E <- I - T^2 + E T^2 (T^2 is in the T buffer); ||E||_F on the way
      double r = 0.0;
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const int d = dg(n);
          double e0 = (d == 1 ? 1.0 : 0.0), e1 = (d == 2 ? 1.0 : 0.0);
          if (op == OP_ET) { const double2 t2 = own(Tp, n); e0 = (e0 - t2.x) + acc[n][0]; e1 = (e1 - t2.y) + acc[n][1]; }
          else { e0 -= acc[n][0]; e1 -= acc[n][1]; }
          const double w = (st.ti[n] != st.tj[n]) ? 2.0 : 1.0;
          r = fma(w * e0, e0, fma(w * e1, e1, r));
          acc[n][0] = e0; acc[n][1] = e1;
        }
      r = sqrt(nsp_block_reduce<NTH>(r, false, red));           // (its barriers: everyone is done reading Y)
      nsp_store<NTW>(ys, nt, st, L, acc);
      __syncthreads();
      op = plan(r);
      if (op < 0) break;
    } else if (op == OP_E2) {
      if (nss_steps[step].kind % 10 == 2) {                      // T = c0 I + c1 E + c2 E^2 (the T buffer is free)
        const double c0 = nss_steps[step].c[0], c1 = nss_steps[step].c[1], c2 = nss_steps[step].c[2];
#pragma unroll
        for (int n = 0; n < NTW; ++n)
          if (n < st.n) {
            const double2 ev = own(Yp, n);
            const int d = dg(n);
            sts_f64x2(ts + cb(n), fma(c2, acc[n][0], fma(c1, ev.x, d == 1 ? c0 : 0.0)),
                      fma(c2, acc[n][1], fma(c1, ev.y, d == 2 ? c0 : 0.0)));
          }
        __syncthreads();
        op = OP_ZT;
      } else {
        nsp_store<NTW>(ts, nt, st, L, acc);                     // E^2 -> T buffer
        __syncthreads();
        op = OP_E3;
      }
    } else if (op == OP_E3) {
      const double c0 = nss_steps[step].c[0], c1 = nss_steps[step].c[1], c2 = nss_steps[step].c[2], c3 = nss_steps[step].c[3];
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const double2 ev = own(Yp, n), e2 = own(Tp, n);
          const int d = dg(n);
          acc[n][0] = fma(c3, acc[n][0], fma(c2, e2.x, fma(c1, ev.x, d == 1 ? c0 : 0.0)));
          acc[n][1] = fma(c3, acc[n][1], fma(c2, e2.y, fma(c1, ev.y, d == 2 ? c0 : 0.0)));
        }
      __syncthreads();                                           // everyone is done reading E^2
      nsp_store<NTW>(ts, nt, st, L, acc);
      __syncthreads();
      op = OP_ZT;
    } else if (op == OP_ZT) {
      __syncthreads();                                           // everyone is done reading Z
      nsp_store<NTW>(zs, nt, st, L, acc);
      if (nss_steps[step].kind > 10) { __syncthreads(); rc = nprod; break; }
      op = OP_T2;                                                // (T^2 reads the T buffer only)
    } else {  // OP_T2
      __syncthreads();                                           // everyone is done reading T (and Z is published)
      nsp_store<NTW>(ts, nt, st, L, acc);
      __syncthreads();
      op = OP_ET;
    }
    if (nprod >= 64) break;                                      // cannot happen: every stage contracts rho
  }
  NSP_TICK(12);
  return rc;
}

// WORK = true: consume the classifying pass's work list (per-level analyses); a separate instantiation
// because the two extra live values of the list mode push the default kernel into spilling.
template <int NT, int NTH, int MINB, bool WORK, bool EXT = false>
__global__ void __launch_bounds__(NTH, MINB) letkf_nsp_kernel(ColParams P, int lch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = NTH / 32;
  constexpr int NTW = (NT * (NT + 1) / 2 + NW - 1) / NW;       // upper-triangular tiles per warp
  constexpr int NTA = (4 * NT + NW - 1) / NW;                  // update tiles per warp (lch <= 32 levels)
  constexpr int PCH = ((NS_PCH + NW - 1) / NW) * NW;           // staged observation rows per chunk (whole rows per warp)
  constexpr int nt = NT, kp = 8 * NT, ks = kp + 4, msz = NT * (NT + 1) / 2 * 64;   // host: NT == ceil(k / 8)
  const int k = P.k, nz = P.nz;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Zp = reinterpret_cast<double*>(smem_raw);
  double* Yp = Zp + msz;
  double* Tp = Yp + msz;
  double* gvec = Tp + msz;
  double* wa = gvec + kp;
  double* tv = wa + kp;
  double* xm = tv + kp;                                        // [NSP_LCH_MAX]
  double* ml = xm + NSP_LCH_MAX;                               // [NSP_LCH_MAX]
  double* red = ml + NSP_LCH_MAX;                              // [16]
  double* sel_sq = red + 16;                                   // [NS_SELCAP] sqrt(rho / sigma^2)
  double* sel_d = sel_sq + NS_SELCAP;                          // [NS_SELCAP] sqrt(rho / sigma^2) * d
  int* sel_row = reinterpret_cast<int*>(sel_d + NS_SELCAP);    // [NS_SELCAP] obs row
  int* warp_cnt = sel_row + NS_SELCAP;                         // [32]
  int* s_int = warp_cnt + 32;                                  // [4]
  double* Ych = Zp;          // [PCH][ks] staged weighted rows (phase 1: no matrix is live)
  const unsigned zs = (unsigned)__cvta_generic_to_shared(Zp);
  const unsigned ys = (unsigned)__cvta_generic_to_shared(Yp);
  const unsigned ts = (unsigned)__cvta_generic_to_shared(Tp);

  const double km1 = (double)(k - 1);
  const double sW = sqrt(km1);
  const bool per_level = P.radius_v > 0.0;
  const int nxf = per_level ? nz : 1;
  const int R = index_reach<EXT>(P.iv, P.radius);
  constexpr bool work = WORK;
  const long long ncols = work ? (long long)*P.work_count : (P.cols ? P.ncols : (long long)P.own_nx * P.own_ny);
  const NspLane L = nsp_lane(lane);
  const NspTiles<NTW> st = nsp_tiles<NTW, NTH>(nt, warp);

  for (long long ci = blockIdx.x; ci < ncols; ci += gridDim.x) {
    int lx, ly, lt_b = 0;
    if constexpr (work) {
      const long long item = P.work_items[ci], c = item / nxf;
      lt_b = (int)(item - c * nxf);
      lx = (int)(c % P.nx); ly = (int)(c / P.nx);
    } else if (P.cols) { long long c = P.cols[ci]; lx = (int)(c % P.nx); ly = (int)(c / P.nx); }
    else { lx = (int)(ci % P.own_nx); ly = (int)(ci / P.own_nx); }
    int gx = P.gx0 + lx, gy = P.gy0 + ly;
    const long long col = (long long)ly * P.nx + lx;
    index_col_coords<EXT>(P.iv, col, gx, gy);
    double* Xg = P.X + col * nz * k;
    int col_iters = 0;
    long long col_npl = 0;
    bool col_fail = false;
    // the column's state is first touched in phase 3: start pulling it towards L2 now
    for (int e = tid * 16; e < nz * k; e += NTH * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(Xg + e));

    const int lt_e = work ? lt_b + 1 : nxf;
    for (int lt = lt_b; lt < lt_e; ++lt) {
      // ---------------- 1. selection, gather, C += Yw^T Yw on the FP64 tensor path, g += Yw^T dw
#ifdef NSP_PROFILE
      if (threadIdx.x == 0) nsp_prof_ = P.stats;
      const long long nsp_col_t0 = clock64();
#endif
      NSP_T0();
      double cacc[NTW][2];
#pragma unroll
      for (int n = 0; n < NTW; ++n) { cacc[n][0] = 0.0; cacc[n][1] = 0.0; }
      double gacc = 0.0;
      if (tid == 0) s_int[0] = 0;
      __syncthreads();
      int npl = 0;
      bool deferred = false;
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
              sd = sq * P.d[orow];
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
            for (int w = 0; w < NW; ++w) tot += warp_cnt[w];
            s_int[0] += tot;
          }
          rb += NTH;
          if (rb >= re) {
            ++cy;
            rows_left = (cy <= cy1);
            if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
          }
          __syncthreads();
        }
        const int nsel = s_int[0];
        if (have_batch && rows_left && nsel + min(NTH, re - rb) <= NS_SELCAP) continue;   // the next batch still fits
#ifndef NSP_NO_DEFER
        if (P.small_items && !rows_left && npl == 0 && nsel > 0 && nsel <= SP_PMAX && 2 * nsel <= k &&
            !(P.W_out && P.w_col == col)) {
          // few local observations: the observation-space kernel does this transform (letkf_smallp.cuh)
          __syncthreads();                                       // everyone has read the count
          if (tid == 0) {
            const unsigned slot = atomicAdd(P.small_count, 1u);
            P.small_items[slot] = col * nxf + lt;
          }
          npl = nsel;
          deferred = true;
          break;
        }
#endif
        NSP_TICK(8);
        for (int c0 = 0; c0 < nsel; c0 += PCH) {
          const int rows = min(PCH, nsel - c0), rows4 = (rows + 3) & ~3;
          // gather: warp w stages rows w, w + NW, ...; all loads issued before the stores
          {
            double v[PCH / NW][(NT * 8 + 31) / 32];
#pragma unroll
            for (int q = 0; q < PCH / NW; ++q) {
              const int r = warp + NW * q;
              const double* src = P.Yp + (long long)sel_row[c0 + min(r, rows - 1)] * k;
#pragma unroll
              for (int jj = 0; jj < (NT * 8 + 31) / 32; ++jj) {
                const int j = lane + 32 * jj;
                v[q][jj] = (r < rows && j < k) ? src[j] : 0.0;
              }
            }
#pragma unroll
            for (int q = 0; q < PCH / NW; ++q) {
              const int r = warp + NW * q;
              if (r < rows4) {
                const double sq = (r < rows) ? sel_sq[c0 + r] : 0.0;
#pragma unroll
                for (int jj = 0; jj < (NT * 8 + 31) / 32; ++jj) {
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
              for (int n = 0; n < NTW; ++n) NSP_DMMA(cacc[n], a[n], b[n]);
            }
            if (tid < k) {
              for (int r = 0; r < rows; ++r) gacc = fma(Ych[r * ks + tid], sel_d[c0 + r], gacc);
            }
          }
          __syncthreads();
        }
        NSP_TICK(9);
        npl += nsel;
        if (tid == 0 && nsel) s_int[0] = 0;      // (nsel == 0: already 0, and no barrier since it was read)
        __syncthreads();
        if (!rows_left) break;
      }
      if (lt == lt_b) col_npl = npl;
      if (deferred) continue;

      // ---------------- 2. Z = A^{-1/2}, A = shift I + C, by coupled Newton-Schulz
      bool ok = true;
      if (npl > 0) {
        // spectrum(A) lies in [shift, shift + ||C||_F]: C is PSD with smallest eigenvalue 0 (Y' 1 = 0)
        const double shift = km1 / P.inflation;
        double fro = 0.0;
#pragma unroll
        for (int n = 0; n < NTW; ++n)
          if (n < st.n) {
            const int i = st.ti[n] * 8 + g, j = st.tj[n] * 8 + 2 * t;
            const double w = (st.ti[n] != st.tj[n]) ? 2.0 : 1.0;
            if (i < k && j < k) fro = fma(w * cacc[n][0], cacc[n][0], fro);
            if (i < k && j + 1 < k) fro = fma(w * cacc[n][1], cacc[n][1], fro);
          }
        fro = sqrt(nsp_block_reduce<NTH>(fro, false, red));
        // (||C||_F is typically 2 - 3x the Schatten-4 bound the iteration works with: a transform far beyond the
        // limit is handed over without spending the A^2 product on it)
        if (!((shift + fro) < 8.0 * NSP_KAPPA_MAX * shift)) {
          if (tid == 0) {
            const unsigned slot = atomicAdd(P.redo_count, 1u);
            P.redo_items[slot] = col * nxf + lt;
          }
          continue;
        }
        // A -> T (shift I on the zero padding)
#pragma unroll
        for (int n = 0; n < NTW; ++n)
          if (n < st.n) {
            const int i = st.ti[n] * 8 + g, j = st.tj[n] * 8 + 2 * t;
            const double d0 = (i == j ? shift : 0.0), d1 = (i == j + 1 ? shift : 0.0);
            const double y0 = (i < k && j < k) ? cacc[n][0] + d0 : d0;
            const double y1 = (i < k && j + 1 < k) ? cacc[n][1] + d1 : d1;
            sts_f64x2(ts + nsp_cbase(st.ti[n], st.tj[n], nt, L), y0, y1);
          }
        if (tid < k) gvec[tid] = gacc;
        __syncthreads();
        NSP_TICK(10);
        const int it = nsp_inverse_sqrt<NT, NTH>(Zp, shift, fro, k);   // products used
#ifdef NSP_PROFILE
        nsp_t_ = clock64();
#endif
        if (it == -2) {
          // condition bound beyond NSP_KAPPA_MAX: symmetric tiles are not trusted there, the full-product
          // kernel (k <= 80) or the Jacobi kernel redoes this transform
          if (tid == 0) {
            const unsigned slot = atomicAdd(P.redo_count, 1u);
            P.redo_items[slot] = col * nxf + lt;
          }
          continue;
        }
        if (it < 0) ok = false;
        col_iters = max(col_iters, it);
        // w = Z (Z g): warp per tile row, lanes read Z in the fragment pattern (row g, columns t, 4 + t
        // of every tile) and reduce over t
        if (ok) {
          for (int pass = 0; pass < 2; ++pass) {
            const double* vin = pass ? tv : gvec;
            double* vout = pass ? wa : tv;
            for (int I = warp; I < nt; I += NW) {
              NspWalk zw;
              zw.start(I);
              double s = 0.0;
#pragma unroll 1
              for (int K = 0; K < nt; ++K) {
                const double z0 = lds_f64(zw.addr(zs, K, 0, L)), z1 = lds_f64(zw.addr(zs, K, 1, L));
                const int c = K * 8 + t;
                s = fma(z0, c < k ? vin[c] : 0.0, s);
                s = fma(z1, c + 4 < k ? vin[c + 4] : 0.0, s);
                zw.next(K, nt);
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              if (t == 0 && I * 8 + g < k) vout[I * 8 + g] = s;
            }
            __syncthreads();
          }
        }
      }
      NSP_TICK(13);
      if (!ok) col_fail = true;

      if (P.W_out && P.w_col == col && lt == 0) {
        for (int e = tid; e < k * k; e += NTH) {
          const int j = e / k, i = e - j * k;
          double vv;
          if (npl == 0) vv = (i == j) ? sqrt(P.inflation) : 0.0;
          else if (!ok) vv = nan("");
          else vv = wa[j] + sW * Zp[nsp_elem(j, i, nt)];
          P.W_out[e] = vv;
        }
        __syncthreads();
      }

      // ---------------- 3. X_a = xbar + X' w + sW X' Z on the tensor path, level chunks of lch,
      //                     staged in the Y and T buffers
      double* Xt = Yp;                                                 // [lch][ks]
      double* To = Xt + (size_t)lch * ks;                              // [lch][k]
      const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
      if (ok) {
        for (int l0 = lev_b; l0 < lev_e; l0 += lch) {
          const int nl = min(lch, lev_e - l0);
          for (int e = tid; e < nl * k; e += NTH) {
            const int l = e / k, j = e - l * k;
            Xt[l * ks + j] = Xg[(long long)l0 * k + e];
          }
          if (kp > k) for (int e = tid; e < nl * (kp - k); e += NTH) Xt[(e / (kp - k)) * ks + k + e % (kp - k)] = 0.0;
          __syncthreads();
          for (int l = warp; l < nl; l += NW) {
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
            for (int e = tid; e < nl * k; e += NTH) { const int l = e / k; To[e] = xm[l] + Xt[l * ks + (e - l * k)] * f; }
          } else {
            const int ntr = (nl + 7) >> 3;
            const NsTiles<NTA> at = nsp_rect_tiles<NTA, NTH>(ntr, nt, warp);
            double uacc[NTA][2];
#pragma unroll
            for (int n = 0; n < NTA; ++n) { uacc[n][0] = 0.0; uacc[n][1] = 0.0; }
            const double* xa = Xt + g * ks + t;
            NspWalk zw[NTA];
#pragma unroll
            for (int n = 0; n < NTA; ++n) zw[n].start(at.tj[n]);
#pragma unroll 1
            for (int K = 0; K < nt; ++K) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                double a[NTA], b[NTA];
#pragma unroll
                for (int n = 0; n < NTA; ++n) {
                  if (n == 0 || at.ti[n] != at.ti[n - 1]) a[n] = xa[at.ti[n] * 8 * ks + K * 8 + h * 4]; else a[n] = a[n - 1];
                  b[n] = lds_f64(zw[n].addr(zs, K, h, L));
                }
#pragma unroll
                for (int n = 0; n < NTA; ++n) NSP_DMMA(uacc[n], a[n], b[n]);
              }
#pragma unroll
              for (int n = 0; n < NTA; ++n) zw[n].next(K, nt);
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
          for (int e = tid; e < nl * k; e += NTH) Xg[(long long)l0 * k + e] = To[e];
          if (P.mean_out) {
            for (int l = warp; l < nl; l += NW) {
              double s = 0.0;
              for (int j = lane; j < k; j += 32) s += To[l * k + j];
              s = warp_sum(s);
              if (lane == 0) P.mean_out[col * nz + l0 + l] = s * (1.0 / (double)k);
            }
          }
          __syncthreads();
        }
      }
      NSP_TICK(14);
#ifdef NSP_PROFILE
      if (threadIdx.x == 0) atomicAdd((unsigned long long*)&P.stats[15], (unsigned long long)(clock64() - nsp_col_t0));
#endif
    }  // lt
    if (tid == 0) {
      if (!work) {                       // (work mode: the classifying pass counted the columns)
        atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)col_npl);
        atomicMax(&P.stats[1], col_npl);
        atomicAdd((unsigned long long*)&P.stats[5], 1ull);
      }
      atomicAdd((unsigned long long*)&P.stats[2], (unsigned long long)col_iters);
      atomicMax(&P.stats[3], (long long)col_iters);
      if (col_fail) atomicAdd((unsigned long long*)&P.stats[4], 1ull);
    }
  }
}

// largest level chunk (multiple of 8, <= 32) whose staging fits in the two free matrix buffers
static int nsp_level_chunk(int k, int nz) {
  const int fit = (2 * nsp_ntiles(k) * 64) / (ns_stride(k) + k);
  return std::max(8, std::min(std::min(32, (nz + 7) & ~7), fit & ~7));
}
static size_t nsp_smem_bytes(int k, int lch, int nth) {
  const int pch = ((NS_PCH + nth / 32 - 1) / (nth / 32)) * (nth / 32);
  size_t mats = 3 * (size_t)nsp_ntiles(k) * 64;
  mats = std::max(mats, (size_t)pch * ns_stride(k));                       // phase-1 staging aliases them
  const size_t dbl = mats + 3 * (size_t)ns_kp(k) + 2 * NSP_LCH_MAX + 16 + 2 * NS_SELCAP;
  return dbl * 8 + (size_t)NS_SELCAP * 4 + 32 * 4 + 4 * 4 + 16;
}
