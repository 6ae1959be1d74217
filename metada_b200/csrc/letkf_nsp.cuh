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
// per-level transforms) whose rigorous condition bound exceeds the limit (default 1e5) is appended to a
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
#if defined(NSP_PROFILE) && defined(NSP_PROFILE_WARPS)
#define NSP_T0() do {} while (0)
#define NSP_TICK(slot) do {} while (0)
__device__ long long* nsp_prof_;
#elif defined(NSP_PROFILE)
#define NSP_T0() long long nsp_t_ = clock64()
#define NSP_TICK(slot) do { const long long n_ = clock64(); if (threadIdx.x == 0) atomicAdd((unsigned long long*)&nsp_prof_[slot], (unsigned long long)(n_ - nsp_t_)); nsp_t_ = n_; } while (0)
__device__ long long* nsp_prof_;
#else
#define NSP_T0() do {} while (0)
#define NSP_TICK(slot) do {} while (0)
#endif
// -DNSP_PROFILE -DNSP_PROFILE_UPDATE: slots 10 / 11 / 12 are re-used for the update's product, state wait and level
// mean passes (8 then holds everything before the iteration, 9 the iteration, 13 w, 14 the stores)
// -DNSP_PROFILE -DNSP_PROFILE_EPI: slots 8 / 9 = wait at the first barrier / store + second barrier of the Z T and T T
// epilogues (the Gram phase's ticks are off)
#if defined(NSP_PROFILE) && defined(NSP_PROFILE_EPI)
#define NSP_TICK3(slot) NSP_TICK(slot)
#else
#define NSP_TICK3(slot) do {} while (0)
#endif
// -DNSP_PROFILE -DNSP_PROFILE_EPI -DNSP_PROFILE_EPI2: slots 8 / 9 = the A^2 epilogue (bound, start) / the residual epilogues
// (M0, ET) instead
#if defined(NSP_PROFILE) && defined(NSP_PROFILE_EPI2)
#define NSP_TICK4(slot) NSP_TICK(slot)
#undef NSP_TICK3
#define NSP_TICK3(slot) do {} while (0)
#else
#define NSP_TICK4(slot) do {} while (0)
#endif
#if defined(NSP_PROFILE) && defined(NSP_PROFILE_UPDATE)
#define NSP_TICK2(slot) NSP_TICK(slot)
#define NSP_SLOT_PRE(slot) 8
#define NSP_SLOT_IT(slot) 9
#else
#define NSP_SLOT_IT(slot) slot
#define NSP_TICK2(slot) do {} while (0)
#if defined(NSP_PROFILE_EPI)
#define NSP_SLOT_PRE(slot) 10
#else
#define NSP_SLOT_PRE(slot) slot
#endif
#endif

__host__ __device__ inline int nsp_ntiles(int k) { const int nt = (k + 7) >> 3; return nt * (nt + 1) / 2; }
// doubles of the staging region: three packed matrices, or NSP_GCH_OF(nt) gathered rows, or one matrix + lch staged levels
__host__ __device__ inline int nsp_region_doubles(int k, int lch) {
  const int kp = (k + 7) & ~7, ks = kp + 4, m = nsp_ntiles(k) * 64;
  int r = 3 * m;
  const int gch = (kp >> 3) <= 5 ? 112 : 128;                  // NSP_GCH_OF
  if (gch * ks > r) r = gch * ks;
  if (m + lch * ks > r) r = m + lch * ks;
  return r;
}
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

// Block sum with ONE barrier: the per-warp partials go to one of two buffers in turn (rbuf flips on every call), so
// the barrier that publishes the partials is also the only one needed -- a buffer is rewritten two reductions later,
// and every thread has passed the barrier in between.  The caller may rely on that barrier for its own hazards.
template <int NTH>
__device__ __forceinline__ double nsp_block_sum1(double v, double* red /*[2][16]*/, int& rbuf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  double* rb = red + 16 * rbuf;
  rbuf ^= 1;
  if (lane == 0) rb[warp] = v;
  __syncthreads();
  double r = rb[0];
#pragma unroll
  for (int w = 1; w < NTH / 32; ++w) r += rb[w];
  return r;
}

// ---- bulk asynchronous copies (the TMA unit, 1-D form) and the mbarrier that tracks their bytes.
// Rows of Y' (k doubles, contiguous in HBM) and the levels of the column's state block go global -> shared without
// passing through registers, all of a column's rows in flight at once; the analysed block goes back with one
// bulk store.  Requirements: 16-byte aligned addresses and sizes (k even).
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NSP_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NSP_DONE_%=;\n"
      "bra NSP_WAIT_%=;\n"
      "NSP_DONE_%=:\n"
      "}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(__cvta_generic_to_global(dst)), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

// The same run of tiles, compactly (five registers instead of fifteen: the iteration keeps it alive across its
// products, whose software pipeline needs the registers): tiles are contiguous in packed storage, so tile m of the
// run sits at (e0 + m) * 512; its coordinates follow by walking from (ti0, tj0).
struct NspRun {
  int e0, n, ti0, tj0;
  unsigned dmask;        // bit m: tile m is on the diagonal
};
template <int NTW, int NTH>
__device__ __forceinline__ NspRun nsp_run(int nt, int warp) {
  const int E = nt * (nt + 1) / 2;
  NspRun r;
  r.e0 = (E * warp) / (NTH / 32);
  r.n = (E * (warp + 1)) / (NTH / 32) - r.e0;
  int i = 0, rowstart = 0;
  while (r.e0 >= rowstart + (nt - i)) { rowstart += nt - i; ++i; }
  r.ti0 = i;
  r.tj0 = i + (r.e0 - rowstart);
  r.dmask = 0u;
  int ti = r.ti0, tj = r.tj0;
#pragma unroll
  for (int m = 0; m < NTW; ++m) {
    if (m < r.n && ti == tj) r.dmask |= 1u << m;
    if (++tj == nt) { ++ti; tj = ti; }
  }
  return r;
}
// thread index the compiler cannot hoist out of a loop (keeps per-warp tile tables from staying live across phases)
__device__ __forceinline__ int nsp_tid_opaque() {
  int v;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v));
  return v;
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
//
// The body is an explicit software pipeline written with VOLATILE loads, so that ptxas keeps its order: left to
// itself ptxas sinks every fragment load to just before the MMA that consumes it and funnels all of them through
// one register (LDS -> DMMA -> LDS ..., each MMA waiting a shared-memory latency): the warps were latency bound in
// proportion to their load count (60 - 90 per k-half), the lightest waited ~40 % of the product at the barrier
// (ncu source view, r02).  Here B fragments are requested two MMAs ahead (ring of three registers) and the A
// fragments of the next k-tile during the current one.
template <int IMM>
__device__ __forceinline__ double nsp_lds_v(unsigned base) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(base), "n"(IMM));
  return v;
}
// fragment element S[8 I + g][8 K + 4 h + t]: d = matrix + straight lane offset of this k-half,
// tr = matrix + transposed lane offset of this k-half (shared-memory byte addresses)
template <int NT, int I, int K>
__device__ __forceinline__ double nsp_frag_v(unsigned d, unsigned tr) {
  if constexpr (I <= K) return nsp_lds_v<(nsp_rs(I, NT) + K - I) * 512>(d);
  else return nsp_lds_v<(nsp_rs(K, NT) + I - K) * 512>(tr);
}

// B fragments are requested NSP_BAHEAD MMAs ahead of their use (ring of NSP_BAHEAD + 1 registers)
// (A/B on one box, r02: 2 ahead 25.25 ms, 3 ahead 25.12, 4 / 5 ahead 25.45 on the C5 probe)
#ifndef NSP_BAHEAD
#define NSP_BAHEAD 3
#endif
#define NSP_BRING (NSP_BAHEAD + 1)
template <int NT, int NW, int NTW, int W>
__device__ __forceinline__ void nsp_mm_warp(unsigned pbase, unsigned qbase, const NspLane& L, double (&acc)[NTW][2]) {
  constexpr int E = NT * (NT + 1) / 2, e0 = (E * W) / NW, e1 = (E * (W + 1)) / NW, N = e1 - e0;
  if constexpr (N > 0) {
    constexpr int R0 = nsp_tile_i(NT, e0), NR = nsp_tile_i(NT, e1 - 1) - R0 + 1;   // the warp's tile rows (NR <= N)
    constexpr int TOT = NT * N;                                                     // MMAs per k-half
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const unsigned pd = pbase + L.offd + (h ? L.dh : 0u), pt = pbase + L.offt + ((unsigned)h << 8);
      const unsigned qd = qbase + L.offd + (h ? L.dh : 0u), qt = qbase + L.offt + ((unsigned)h << 8);
      double ar[2][NR], br[NSP_BRING];
      nsp_static_for<0, NR>([&](auto r_c) {
        constexpr int r = decltype(r_c)::value;
        ar[0][r] = nsp_frag_v<NT, R0 + r, 0>(pd, pt);
      });
      nsp_static_for<0, NSP_BAHEAD>([&](auto j_c) {
        constexpr int j = decltype(j_c)::value;
        if constexpr (j < TOT) br[j] = nsp_frag_v<NT, nsp_tile_j(NT, e0 + j % N), j / N>(qd, qt);
      });
      nsp_static_for<0, TOT>([&](auto i_c) {
        constexpr int idx = decltype(i_c)::value, K = idx / N, n = idx % N;
        if constexpr (idx + NSP_BAHEAD < TOT) {
          constexpr int K2 = (idx + NSP_BAHEAD) / N, n2 = (idx + NSP_BAHEAD) % N;
          br[(idx + NSP_BAHEAD) % NSP_BRING] = nsp_frag_v<NT, nsp_tile_j(NT, e0 + n2), K2>(qd, qt);
        }
        if constexpr (K + 1 < NT && n < NR) ar[(K + 1) & 1][n] = nsp_frag_v<NT, R0 + n, K + 1>(pd, pt);
        NSP_DMMA(acc[n], ar[K & 1][nsp_tile_i(NT, e0 + n) - R0], br[idx % NSP_BRING]);
      });
    }
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
__device__ __forceinline__ void nsp_mm_any(unsigned pbase, unsigned qbase, int warp, const NspLane& L, double (&acc)[NTW][2]) {
#pragma unroll
  for (int n = 0; n < NTW; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
  if constexpr (NT <= NSP_FIXED_MAX_NT) {
    nsp_mm_dispatch<NT, NW, NTW>(std::make_integer_sequence<int, NW>{}, warp, pbase, qbase, L, acc);
  } else {
    const NspTiles<NTW> w = nsp_tiles<NTW, NW * 32>(NT, warp);
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

// the run's accumulator pairs -> packed storage at dst0 = matrix + (e0 << 9) + lane offset
template <int NTW>
__device__ __forceinline__ void nsp_store_run(unsigned dst0, int n_tiles, double (&acc)[NTW][2]) {
#pragma unroll
  for (int n = 0; n < NTW; ++n)
    if (n < n_tiles) sts_f64x2(dst0 + ((unsigned)n << 9), acc[n][0], acc[n][1]);
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

// ---- schedule look-up (ns_schedule_table.h): all threads compute the same index (constant-memory broadcasts)
__device__ __forceinline__ int nss_start_index(double kappa) {
  // kappa grid: kappa_i - 1 = 1e-3 * 1.12^i; smallest i with kappa_i >= kappa
  int i = (int)ceilf(__log2f(fmaxf((float)(kappa - 1.0), 1e-3f) * 1e3f) * (1.0f / 0.16349873f));
  i = max(0, min(i, NSS_NKAPPA - 1));
  while (i > 0 && nss_starts[i - 1].kappa >= kappa) --i;
  while (i < NSS_NKAPPA - 1 && nss_starts[i].kappa < kappa) ++i;
  return (nss_starts[i].kappa >= kappa) ? i : -1;
}

// largest condition bound (Schatten-4 bound of the spectrum / shift) the packed kernel takes by default
// (mdc_letkf_params.kappa_max overrides it, up to NSP_KAPPA_TABLE_MAX); beyond it the transform goes to the redo
// list.  Symmetric-tile products assume the iterates commute; with the short composite minimax schedule (13 - 25
// products) the rounding defect stays at the level of the eigen-decomposition's own: difference of Z to numpy's eigh
// (emulation of these very tile products, tests/ns_emul.py) 5e-15 at cond 50, 3e-14 at 1200, 1.5e-13 at 6e3, 6e-13 at
// 5e4, <= 8e-12 at 1e5 over k = 24 .. 128 -- and the residual Z A Z - I in long double is within 2x of that of the
// eigen-decomposition at every one of them (round 1 stopped at 256, round 2's first table at 2000: the accurate-
// observation cliff of bench.py's sigma = 0.01 variant).  The MEAN update w = Z (Z g) alone loses err(Z) sqrt(cond)
// (2e-10 at cond 7e4) and gets one step of iterative refinement beyond NSP_REFINE_KAPPA (nsp_phase_update); with it
// the analysis agrees with the oracle's eigen-decomposition to 4e-12 at cond 7e4, 9e-12 at 1.7e5 and 1.1e-11 at the
// table's end (tools/refine_probe.py on the device, k = 40 ... 128), so the default limit leaves a factor 20.
#define NSP_KAPPA_MAX_DEFAULT 1e5
#define NSP_KAPPA_TABLE_MAX 3e5
#define NSP_SC_DOUBLES 48   /* schedule scratch: 8 steps x {kind, c0..c3}, + the residual bound of the finish */
#define NSP_SC_KMAX 45      /* slot that carries the condition limit from the Gram phase to the iteration */
#define NSP_SC_ROUNDS 44    /* selection rounds of the Gram phase (1: sel_row / sel_w still describe every local observation) */
#define NSP_SC_KAPPA 43     /* the transform's condition bound, from the iteration to the update */
#define NSP_REFINE_KAPPA 1e4      /* beyond it the mean update gets one step of iterative refinement (below: <= 1e-11 without) */
#define NSP_RCH 64          /* rows re-gathered per round of the refinement */

// Z <- A^{-1/2} for the A held in the T buffer, by the composite minimax polynomial iteration of
// tools/gen_ns_schedule.py: state Z and the residual E = I - Z^2 A (in the Y buffer), spectrum(E) in [-rho, rho];
//   start  (degree 0..2 in A, chosen by the condition bound kappa):  Z0 = q(A), E0 = I - Z0 (A Z0)
//   stage  (degree d = 1..3):  T = sum c_i E^i,  Z <- Z T,  E <- I - T^2 + E T^2       d + 2 products
//   finish (degree f = 1..3):  Z <- Z sum c_i E^i                                       f products
// with the minimax coefficients for the interval the spectrum is known to lie in and the degree sequence that
// minimises the product count (13 - 14 products at C5's conditioning; Chebyshev start + Newton-Schulz + series
// finish took 17 - 20).  Every rho follows a priori from kappa (rigorous while spectrum(A) is inside
// [shift, shift kappa]), so the start entry lists the column's whole sequence of steps: their coefficients are
// copied to shared memory once (sc), and no epilogue waits on a table look-up; ||E||_F is measured once, before the
// finishing step, against sqrt(k) rho.
// Inlined into the kernel body, which is kept thin (the other phases are separate functions): ptxas schedules the
// products' software pipeline as written only when few registers are alive around it -- inside a fat column loop,
// and also as a called function of one, it fell back to a pressure-minimising schedule that serialised every
// fragment load behind the MMA before it (r02 SASS; the isolated function pipelines even with 80 registers).
// Returns the number of k x k products used, -1 if the iteration failed, -2 if kappa is beyond the limit (sc[NSP_SC_KMAX]).
#ifdef NSP_ISQ_NOINLINE   /* development */
#define NSP_ISQ_INLINE __noinline__
#else
#define NSP_ISQ_INLINE __forceinline__
#endif
template <int NT, int NTH>
__device__ NSP_ISQ_INLINE int nsp_inverse_sqrt(double* Zp, double* red, double* sc, int& rbuf, double shift, double rs,
                                             double fro, int k) {
  constexpr int NW = NTH / 32, NTW = (NT * (NT + 1) / 2 + NW - 1) / NW, nt = NT, kp = 8 * NT;
  constexpr int msz = NT * (NT + 1) / 2 * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Yp = Zp + msz;
  double* Tp = Yp + msz;
  const unsigned zs = (unsigned)__cvta_generic_to_shared(Zp), ys = zs + msz * 8, ts = ys + msz * 8;
  const NspLane L = nsp_lane(lane);
  const NspRun st = nsp_run<NTW, NTH>(nt, warp);
  // the lane's accumulator pair of tile n sits at byte offset cb(n) of every packed matrix; dg(n): 0 = off the
  // diagonal, 1 = first element on it, 2 = second
  const unsigned cb0 = ((unsigned)st.e0 << 9) + L.offc;
  const int ldg = (g == 2 * t) ? 1 : (g == 2 * t + 1 ? 2 : 0);
  auto cb = [&](int n) { return cb0 + ((unsigned)n << 9); };
  auto dg = [&](int n) { return ((st.dmask >> n) & 1u) ? ldg : 0; };
  auto offw = [&](int n) { return ((st.dmask >> n) & 1u) ? 1.0 : 2.0; };   // weight of the tile in a Frobenius norm
  auto own = [&](const double* M, int n) { return *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(M) + cb(n)); };
  // The iteration is a small state machine around the SINGLE product site (nsp_mm_any; its per-warp specialised
  // bodies must exist once for the instruction cache).  Operands by op:
  enum { OP_A2, OP_Y0, OP_M0, OP_E2, OP_E3, OP_ZT, OP_T2, OP_ET };
  double acc[NTW][2];
  int op = OP_A2, nprod = 0, sn = -1;       // sn: position of the current step in the column's sequence
  int rc = -1;
  // T = c0 I + c1 E (E in acc) for a degree-1 step
  auto store_T_lin = [&](const double* c) {
    const double c0 = c[1], c1 = c[2];
#pragma unroll
    for (int n = 0; n < NTW; ++n)
      if (n < st.n) {
        const int d = dg(n);
        sts_f64x2(ts + cb(n), fma(c1, acc[n][0], d == 1 ? c0 : 0.0), fma(c1, acc[n][1], d == 2 ? c0 : 0.0));
      }
  };
  NSP_T0();
#pragma unroll 1
  while (true) {
    const unsigned pb = (op == OP_A2 || op == OP_Y0 || op == OP_T2) ? ts : (op == OP_M0 || op == OP_ZT) ? zs : ys;
    const unsigned qb = (op == OP_Y0) ? zs : (op == OP_M0 || op == OP_E2) ? ys : ts;
    NSP_TICK(NSP_SLOT_IT(12));
#ifdef NSP_PROFILE_WARPS   /* development: stats[8 + warp] = cycles warp `warp` spends inside the products (warps 0..7) */
    const long long pw_t0_ = clock64();
#endif
    nsp_mm_any<NT, NW, NTW>(pb, qb, warp, L, acc);
#ifdef NSP_PROFILE_WARPS
    if (lane == 0 && warp < 8) atomicAdd((unsigned long long*)&nsp_prof_[8 + warp], (unsigned long long)(clock64() - pw_t0_));
#endif
    NSP_TICK(NSP_SLOT_IT(11));
    ++nprod;
    if (op == OP_A2) {
      // tighter upper end of the spectrum from the product just made: lmax(C)^2 <= ||C^2||_F
      // (Schatten-4 norm of C; C^2 = A^2 - 2 shift A + shift^2 I), typically 2-3x below ||C||_F
      double f4 = 0.0;
      int wti = st.ti0, wtj = st.tj0;
#pragma unroll
      for (int n = 0; n < NTW; ++n) {
        if (n < st.n) {
          const int i = wti * 8 + g, j = wtj * 8 + 2 * t, d = dg(n);
          const double2 a = own(Tp, n);
          const double w = offw(n);
          const double q0 = acc[n][0] - 2.0 * shift * a.x + (d == 1 ? shift * shift : 0.0);
          const double q1 = acc[n][1] - 2.0 * shift * a.y + (d == 2 ? shift * shift : 0.0);
          if (i < k && j < k) f4 = fma(w * q0, q0, f4);
          if (i < k && j + 1 < k) f4 = fma(w * q1, q1, f4);
        }
        if (++wtj == nt) { ++wti; wtj = wti; }
      }
      f4 = nsp_block_sum1<NTH>(f4, red, rbuf);                   // (its barrier: every warp is done with A A)
      // kappa in single precision, rounded up (a bound that selects a table entry): no FP64 sqrt / division chain
      const float hi_f = (float)(shift + fro) * 1.000001f, sh_f = (float)shift;
      const float s4 = sqrtf(sqrtf((float)f4 * 1.000001f) + 1e-13f * hi_f * hi_f) * 1.000001f;
      const double kappa = (double)(fmaxf(fminf(hi_f, sh_f * 1.000001f + s4) / (sh_f * 0.999999f), 1.0f) * 1.000002f);
      const int si = (kappa <= sc[NSP_SC_KMAX]) ? nss_start_index(kappa) : -1;
      if (si < 0) { rc = -2; break; }
      if (tid == 0) sc[NSP_SC_KAPPA] = kappa;
      const int sdeg = nss_starts[si].degree;
      // the column's steps -> shared memory: thread e copies {kind, c0..c3}[e % 5] of step e / 5
      if (tid < 40) {
        const int q = tid / 5, c = tid - 5 * q;
        double v = 0.0;
        if (q < nss_starts[si].nsteps) {
          const int j = nss_starts[si].step[q];
          v = (c == 0) ? (double)nss_steps[j].kind : nss_steps[j].c[c - 1];
          if (c == 0 && nss_steps[j].kind > 10) sc[40] = nss_steps[j].rho;     // residual bound of the finish
        }
        sc[tid] = v;
      }
      const double is = rs * rs;
      const double z0 = nss_starts[si].a[0] * rs, z1 = nss_starts[si].a[1] * rs * is, z2 = nss_starts[si].a[2] * rs * is * is;
      // Z0 = q(A) -> Z buffer; degree 0: E0 = I - z0^2 A, degree 1: Y0 = A Z0 = z0 A + z1 A^2 -> Y buffer (no product)
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const double2 a = own(Tp, n);
          const int d = dg(n);
          const double dx = (d == 1 ? 1.0 : 0.0), dy = (d == 2 ? 1.0 : 0.0);
          sts_f64x2(zs + cb(n), fma(z2, acc[n][0], fma(z1, a.x, z0 * dx)), fma(z2, acc[n][1], fma(z1, a.y, z0 * dy)));
          if (sdeg == 0) {
            acc[n][0] = dx - z0 * z0 * a.x; acc[n][1] = dy - z0 * z0 * a.y;
            sts_f64x2(ys + cb(n), acc[n][0], acc[n][1]);
          } else if (sdeg == 1) {
            sts_f64x2(ys + cb(n), fma(z1, acc[n][0], z0 * a.x), fma(z1, acc[n][1], z0 * a.y));
          }
        }
      __syncthreads();
      if (sdeg == 0) {
        // E0 is the first residual (not measured: rho0 = (kappa - 1) / (kappa + 1) is exact for the bound)
        sn = 0;
        if ((int)sc[0] % 10 == 1) { store_T_lin(sc); __syncthreads(); op = OP_ZT; }
        else op = OP_E2;
      } else {
        op = (sdeg == 1) ? OP_M0 : OP_Y0;
      }
      NSP_TICK4(8);
    } else if (op == OP_Y0) {
      nsp_store_run<NTW>(ys + cb0, st.n, acc);                       // Y0 = A Z0 (the Y buffer is free)
      __syncthreads();
      op = OP_M0;
    } else if (op == OP_M0 || op == OP_ET) {
      // E = I - Z0 Y0, or E <- I - T^2 + E T^2 (T^2 is in the T buffer)
      ++sn;
      const double* c = sc + 5 * sn;                             // the step that starts now
      const int kind = (int)c[0];
      double r = 0.0;
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const int d = dg(n);
          double e0 = (d == 1 ? 1.0 : 0.0), e1 = (d == 2 ? 1.0 : 0.0);
          if (op == OP_ET) { const double2 t2 = own(Tp, n); e0 = (e0 - t2.x) + acc[n][0]; e1 = (e1 - t2.y) + acc[n][1]; }
          else { e0 -= acc[n][0]; e1 -= acc[n][1]; }
          if (kind > 10) {                                       // (the residual is measured once, before the finish)
            const double w = offw(n);
            r = fma(w * e0, e0, fma(w * e1, e1, r));
          }
          acc[n][0] = e0; acc[n][1] = e1;
        }
      if (kind > 10) {
        // last residual: ||E||_F^2 <= k rho^2 must hold if the spectrum is where the bounds say (also catches NaN)
        r = nsp_block_sum1<NTH>(r, red, rbuf);                   // (its barrier: everyone is done reading Y and T)
        const double bound = sc[40] * 1.01;
        if (!(r <= (double)kp * bound * bound)) break;
      } else {
        __syncthreads();                                         // everyone is done reading Y and T
      }
      nsp_store_run<NTW>(ys + cb0, st.n, acc);
      if (kind % 10 == 1) store_T_lin(c);
      __syncthreads();
      op = (kind % 10 == 1) ? OP_ZT : OP_E2;
      NSP_TICK4(9);
    } else if (op == OP_E2) {
      const double* c = sc + 5 * sn;
      if ((int)c[0] % 10 == 2) {                                 // T = c0 I + c1 E + c2 E^2 (the T buffer is free)
        const double c0 = c[1], c1 = c[2], c2 = c[3];
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
        nsp_store_run<NTW>(ts + cb0, st.n, acc);                     // E^2 -> T buffer
        __syncthreads();
        op = OP_E3;
      }
    } else if (op == OP_E3) {
      const double* c = sc + 5 * sn;
      const double c0 = c[1], c1 = c[2], c2 = c[3], c3 = c[4];
#pragma unroll
      for (int n = 0; n < NTW; ++n)
        if (n < st.n) {
          const double2 ev = own(Yp, n), e2 = own(Tp, n);
          const int d = dg(n);
          acc[n][0] = fma(c3, acc[n][0], fma(c2, e2.x, fma(c1, ev.x, d == 1 ? c0 : 0.0)));
          acc[n][1] = fma(c3, acc[n][1], fma(c2, e2.y, fma(c1, ev.y, d == 2 ? c0 : 0.0)));
        }
      __syncthreads();                                           // everyone is done reading E^2
      nsp_store_run<NTW>(ts + cb0, st.n, acc);
      __syncthreads();
      op = OP_ZT;
    } else if (op == OP_ZT) {
      __syncthreads();                                           // everyone is done reading Z
      NSP_TICK3(8);
      nsp_store_run<NTW>(zs + cb0, st.n, acc);
      NSP_TICK3(9);
      if ((int)sc[5 * sn] > 10) { __syncthreads(); rc = nprod; break; }
      op = OP_T2;                                                // (T^2 reads the T buffer only)
    } else {  // OP_T2
      __syncthreads();                                           // everyone is done reading T (and Z is published)
      NSP_TICK3(8);
      nsp_store_run<NTW>(ts + cb0, st.n, acc);
      __syncthreads();
      NSP_TICK3(9);
      op = OP_ET;
    }
    if (nprod >= 64) break;                                      // cannot happen: the sequence ends with a finish
  }
  NSP_TICK(NSP_SLOT_IT(12));
  return rc;
}

// ---- update product U = X' Z for 32 staged levels (4 row tiles) when every warp's tiles lie in one row tile and one
// group of NTA column tiles (NT % NTA == 0, 4 NT % NW == 0; k = 80: 5 tiles, two groups): compile-time tile offsets
// into the packed Z, the same explicit software pipeline as the iteration's products.  xa: shared address of
// X'[row tile of the warp, row g][t]; CG: the warp's column group.
template <int NT, int NTA, int CG>
__device__ __forceinline__ void nsp_upd_warp(unsigned xa, unsigned zbase, const NspLane& L, double (&uacc)[NTA][2]) {
  constexpr int TOT = NT * NTA;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const unsigned zd = zbase + L.offd + (h ? L.dh : 0u), zt = zbase + L.offt + ((unsigned)h << 8);
    const unsigned xh = xa + (unsigned)h * 32u;
    double ar[2], br[3];
    ar[0] = nsp_lds_v<0>(xh);
    br[0] = nsp_frag_v<NT, CG * NTA, 0>(zd, zt);
    if constexpr (TOT > 1) br[1] = nsp_frag_v<NT, CG * NTA + 1 % NTA, 1 / NTA>(zd, zt);
    nsp_static_for<0, TOT>([&](auto i_c) {
      constexpr int idx = decltype(i_c)::value, K = idx / NTA, n = idx % NTA;
      if constexpr (idx + 2 < TOT) {
        constexpr int K2 = (idx + 2) / NTA, n2 = (idx + 2) % NTA;
        br[(idx + 2) % 3] = nsp_frag_v<NT, CG * NTA + n2, K2>(zd, zt);
      }
      if constexpr (K + 1 < NT && n == 0) ar[(K + 1) & 1] = nsp_lds_v<(K + 1) * 64>(xh);
      NSP_DMMA(uacc[n], ar[K & 1], br[idx % 3]);
    });
  }
}
template <int NT, int NTA, int... CGs>
__device__ __forceinline__ void nsp_upd_dispatch(std::integer_sequence<int, CGs...>, int cg, unsigned xa, unsigned zbase,
                                                 const NspLane& L, double (&uacc)[NTA][2]) {
  ((cg == CGs ? (nsp_upd_warp<NT, NTA, CGs>(xa, zbase, L, uacc), 0) : 0), ...);
}

// (the generic tile walk of the larger ensembles keeps many registers of its own: there the iteration is better
// off as a called function)
template <int NT, int NTH>
__device__ __noinline__ int nsp_inverse_sqrt_call(double* Zp, double* red, double* sc, int& rbuf, double shift, double rs,
                                                  double fro, int k) {
  return nsp_inverse_sqrt<NT, NTH>(Zp, red, sc, rbuf, shift, rs, fro, k);
}

/* observation rows staged per gather round: 128, 112 for the small ensembles that run four CTAs per SM */
#define NSP_GCH_OF(NT) ((NT) <= 5 ? 112 : 128)
#define NSP_LCH_CAP 64   /* levels staged per update round */
#define NSP_LSUB 32      /* levels per update product (accumulator tiles per warp) */
#define NSP_SC_FRO 46    /* slot of the schedule scratch that carries ||C||_F from the Gram phase to the iteration */

// Shared memory: one staging region [max(3 matrices, NSP_GCH_OF(nt) rows, matrix + lch levels)] that holds, in turn,
//   phase 1  the gathered Y' rows of the column's local observations (stride k + 4, bulk copies, one per row),
//   phase 2  the three packed matrices Z | Y(E) | T of the iteration,
//   phase 3  Z | the column's state block [lch][k + 4] (bulk copies, one per level), overwritten in place by the
//            analysed block [lch][k] (dense), which one bulk store writes back,
// followed by the small vectors.  Every phase is its own (not inlined) function and rebuilds these pointers: what
// is alive across the call of the iteration decides how many registers its products get -- with the whole column
// loop in one body ptxas had ~55 left for them and serialised every fragment load behind the MMA before it.
template <int NT>
struct NspSm {
  static constexpr int kp = 8 * NT, ks = kp + 4, msz = NT * (NT + 1) / 2 * 64;
  double *Zp, *Yp, *Tp, *gvec, *wa, *tv, *xm, *ml, *red, *sc, *sel_w, *sel_d;
  unsigned long long* mbar_p;   // [0] gather, [1] state block
  int *sel_row, *warp_cnt, *par; // par: [0], [1] parity of the next phase of the two mbarriers
  unsigned zs, ys, ts, mbar_g, mbar_x;
  __device__ __forceinline__ NspSm(int k, int lch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Zp = reinterpret_cast<double*>(smem_raw);
    Yp = Zp + msz;
    Tp = Yp + msz;
    gvec = Zp + nsp_region_doubles(k, lch);
    wa = gvec + kp;
    tv = wa + kp;
    xm = tv + kp;                                              // [NSP_LCH_CAP]
    ml = xm + NSP_LCH_CAP;                                     // [NSP_LCH_CAP]
    red = ml + NSP_LCH_CAP;                                    // [2][16]
    sc = red + 32;                                             // [NSP_SC_DOUBLES] the column's schedule (nsp_inverse_sqrt)
    sel_w = sc + NSP_SC_DOUBLES;                               // [NS_SELCAP] rho / sigma^2
    sel_d = sel_w + NS_SELCAP;                                 // [NS_SELCAP] rho / sigma^2 * d
    mbar_p = reinterpret_cast<unsigned long long*>(sel_d + NS_SELCAP);
    sel_row = reinterpret_cast<int*>(mbar_p + 2);              // [NS_SELCAP] obs row
    warp_cnt = sel_row + NS_SELCAP;                            // [2][32]
    par = warp_cnt + 64;                                       // [4]
    zs = (unsigned)__cvta_generic_to_shared(Zp);
    ys = zs + msz * 8;
    ts = ys + msz * 8;
    mbar_g = (unsigned)__cvta_generic_to_shared(mbar_p);
    mbar_x = mbar_g + 8;
  }
};

// ---------------- phase 1: selection, gather, C = Y^T diag(w) Y on the FP64 tensor path, g = Y^T (w d); A = shift I + C
// goes to the T buffer, g to gvec, ||C||_F to sc[NSP_SC_FRO].
// Returns npl (local observations) | status << 28:  0 go on (npl may be 0: nothing to solve), 1 handed to the
// observation-space kernel, 2 handed to the redo list (condition bound).
template <int NT, int NTH, bool EXT>
__device__ __noinline__ int nsp_phase_gram(const ColParams& P, int lch, long long col, int gx, int gy, int lt) {
  constexpr int NW = NTH / 32;
  constexpr int NTW = (NT * (NT + 1) / 2 + NW - 1) / NW;       // upper-triangular tiles per warp
  constexpr int NGP = (NT * 8 + 31) / 32;                      // members per lane in row-wise passes
  constexpr int nt = NT, kp = 8 * NT, ks = kp + 4;
  const int k = P.k;
  const NspSm<NT> S(k, lch);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Ych = S.Zp;          // [NSP_GCH_OF(NT)][ks] staged rows (no matrix is live)
  double* gpart = S.Zp;        // [NW][kp] per-warp partial sums of g
  const bool tma = (k & 1) == 0;                               // 16-byte aligned rows
  unsigned mph = (unsigned)S.par[0];
  int rbuf = 0, wbuf = 0;
  const bool per_level = P.radius_v > 0.0;
  const int nxf = per_level ? P.nz : 1;
  const int R = index_reach<EXT>(P.iv, P.radius);
  const NspLane L = nsp_lane(lane);
  NSP_T0();
  double cacc[NTW][2];
#pragma unroll
  for (int n = 0; n < NTW; ++n) { cacc[n][0] = 0.0; cacc[n][1] = 0.0; }
  double gp[NGP];
#pragma unroll
  for (int jj = 0; jj < NGP; ++jj) gp[jj] = 0.0;
  int npl = 0, rounds = 0;
  int cy0 = 0, cy1 = -1;
  if (P.radius >= 0.0) index_cy_range(P.iv, gy, R, cy0, cy1);
  int cy = cy0, rb = 0, re = 0;
  bool rows_left = (cy <= cy1);
  if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
  while (true) {
    ++rounds;
    // candidates: the concatenation of the cell rows' index ranges, NTH per batch (a batch may span rows);
    // the selected ones keep that order (the summation order of C is the same whatever the launch shape)
    int nsel = 0;
    while (rows_left && nsel + NTH <= NS_SELCAP) {
      int a = -1, consumed = 0;
      while (rows_left && consumed < NTH) {
        const int take = min(re - rb, NTH - consumed);
        if (tid >= consumed && tid < consumed + take) a = rb + (tid - consumed);
        consumed += take;
        rb += take;
        if (rb >= re) {
          ++cy;
          rows_left = (cy <= cy1);
          if (rows_left) index_row_range(P.iv, gx, R, cy, rb, re);
        }
      }
      bool sel = false;
      double w2 = 0.0, wd = 0.0;
      int orow = 0;
      if (a >= 0) {
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
          w2 = rho * (P.use_R ? ivar : 1.0);
          wd = w2 > 0.0 ? w2 * P.d[orow] : 0.0;            // (a missing value may be NaN: weight 0 drops it)
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, sel);
      int* wc = S.warp_cnt + 32 * wbuf;
      wbuf ^= 1;
      if (lane == 0) wc[warp] = __popc(bal);
      __syncthreads();
      int off = nsel;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int c = wc[w];
        if (w < warp) off += c;
        nsel += c;
      }
      if (sel) {
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        S.sel_row[pos] = orow;
        S.sel_w[pos] = w2;
        S.sel_d[pos] = wd;
      }
    }
    if (tid == 0) bulk_wait_read();                          // the staging is about to be refilled: the previous
    __syncthreads();                                         // column's bulk store has read it; selection published
#ifndef NSP_NO_DEFER
    if (P.small_items && !rows_left && npl == 0 && nsel > 0 && nsel <= SP_PMAX && 2 * nsel <= k &&
        !(P.W_out && P.w_col == col)) {
      // few local observations: the observation-space kernel does this transform (letkf_smallp.cuh)
      if (tid == 0) {
        const unsigned slot = atomicAdd(P.small_count, 1u);
        P.small_items[slot] = col * nxf + lt;
      }
      return nsel | (1 << 28);
    }
#endif
    NSP_TICK(NSP_SLOT_PRE(8));
    for (int c0 = 0; c0 < nsel; c0 += NSP_GCH_OF(NT)) {
      const int rows = min(NSP_GCH_OF(NT), nsel - c0), rows4 = (rows + 3) & ~3;
      // gather: one bulk copy per row, all in flight at once, completion counted in bytes by the mbarrier
      if (tma) {
        fence_proxy_async();                                 // earlier generic accesses of the staging first
        if (tid == 0) mbar_expect_tx(S.mbar_g, (unsigned)(rows * k * 8));
        for (int r = tid; r < rows; r += NTH)
          bulk_g2s(S.zs + (unsigned)(r * ks * 8), P.Yp + (long long)S.sel_row[c0 + r] * k, (unsigned)(k * 8), S.mbar_g);
      } else {
        for (int r = warp; r < rows; r += NW) {
          const double* src = P.Yp + (long long)S.sel_row[c0 + r] * k;
          for (int j = lane; j < k; j += 32) Ych[r * ks + j] = src[j];
        }
      }
      // zero padding: members k..kp-1 of every row, rows up to the next multiple of 4
      if (kp > k)
        for (int e = tid; e < rows * (kp - k); e += NTH) Ych[(e / (kp - k)) * ks + k + e % (kp - k)] = 0.0;
      for (int e = tid; e < (rows4 - rows) * kp; e += NTH) Ych[(rows + e / kp) * ks + e % kp] = 0.0;
      if (tma) { mbar_wait(S.mbar_g, mph); mph ^= 1u; }
      __syncthreads();
      {
        const NspTiles<NTW> st = nsp_tiles<NTW, NTH>(nt, warp);
        const double* ya = Ych + t * ks + g;
        const double* wr = S.sel_w + c0 + t;
        for (int kk = 0; kk < rows4; kk += 4) {
          const double w = (kk + t < rows) ? wr[kk] : 0.0;
          double a[NTW], b[NTW];
#pragma unroll
          for (int n = 0; n < NTW; ++n) {
            if (n == 0 || st.ti[n] != st.ti[n - 1]) a[n] = ya[kk * ks + st.ti[n] * 8] * w; else a[n] = a[n - 1];
            b[n] = ya[kk * ks + st.tj[n] * 8];
          }
#pragma unroll
          for (int n = 0; n < NTW; ++n) NSP_DMMA(cacc[n], a[n], b[n]);
        }
        // g: warp w sums rows w, w + NW, ... (lanes over the members); the warps' partials are added below
        for (int r = warp; r < rows; r += NW) {
          const double dd = S.sel_d[c0 + r];
#pragma unroll
          for (int jj = 0; jj < NGP; ++jj) {
            const int j = lane + 32 * jj;
            if (j < kp) gp[jj] = fma(Ych[r * ks + j], dd, gp[jj]);
          }
        }
      }
      __syncthreads();
    }
    NSP_TICK(NSP_SLOT_PRE(9));
    npl += nsel;
    if (!rows_left) break;
  }
  if (tid == 0) S.par[0] = (int)mph;                         // (published by the barriers below / of the next phase)
  if (npl == 0) { __syncthreads(); return 0; }

  // spectrum(A) lies in [shift, shift + ||C||_F]: C is PSD with smallest eigenvalue 0 (Y' 1 = 0)
  const double shift = (double)(k - 1) / P.inflation;
  const NspRun st = nsp_run<NTW, NTH>(nt, warp);
  const unsigned cb0 = ((unsigned)st.e0 << 9) + L.offc;
  double fro = 0.0;
  {
    int wti = st.ti0, wtj = st.tj0;
#pragma unroll
    for (int n = 0; n < NTW; ++n) {
      if (n < st.n) {
        const int i = wti * 8 + g, j = wtj * 8 + 2 * t;
        const double w = ((st.dmask >> n) & 1u) ? 1.0 : 2.0;
        if (i < k && j < k) fro = fma(w * cacc[n][0], cacc[n][0], fro);
        if (i < k && j + 1 < k) fro = fma(w * cacc[n][1], cacc[n][1], fro);
      }
      if (++wtj == nt) { ++wti; wtj = wti; }
    }
  }
#pragma unroll
  for (int jj = 0; jj < NGP; ++jj) {
    const int j = lane + 32 * jj;
    if (j < kp) gpart[warp * kp + j] = gp[jj];
  }
  fro = sqrt(nsp_block_sum1<NTH>(fro, S.red, rbuf));         // (its barrier publishes the partials of g)
  // (||C||_F is typically 2 - 3x the Schatten-4 bound the iteration works with: a transform far beyond the
  // limit is handed over without spending the A^2 product on it)
  if (!((shift + fro) < 8.0 * P.kappa_max * shift)) {
    if (tid == 0) {
      const unsigned slot = atomicAdd(P.redo_count, 1u);
      P.redo_items[slot] = col * nxf + lt;
    }
    __syncthreads();
    return npl | (2 << 28);
  }
  if (tid < k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += gpart[w * kp + tid];
    S.gvec[tid] = s;
  }
  if (tid == 0) { S.sc[NSP_SC_FRO] = fro; S.sc[NSP_SC_KMAX] = P.kappa_max; S.sc[NSP_SC_ROUNDS] = (double)rounds; }
  // A -> T (shift I on the zero padding)
  {
    int wti = st.ti0, wtj = st.tj0;
#pragma unroll
    for (int n = 0; n < NTW; ++n) {
      if (n < st.n) {
        const int i = wti * 8 + g, j = wtj * 8 + 2 * t;
        const double d0 = (i == j ? shift : 0.0), d1 = (i == j + 1 ? shift : 0.0);
        const double y0 = (i < k && j < k) ? cacc[n][0] + d0 : d0;
        const double y1 = (i < k && j + 1 < k) ? cacc[n][1] + d1 : d1;
        sts_f64x2(S.ts + cb0 + ((unsigned)n << 9), y0, y1);
      }
      if (++wtj == nt) { ++wti; wtj = wti; }
    }
  }
  __syncthreads();
  NSP_TICK(NSP_SLOT_PRE(10));
  return npl;
}

// ---------------- phase 3: X_a = xbar + X' w + sW X' Z on the tensor path (npl = 0: mean kept, perturbations
// inflated).  The state block is requested first (the Y and T buffers are free) and lands while w = Z (Z g) is
// computed.
template <int NT, int NTH>
__device__ __noinline__ void nsp_phase_update(const ColParams& P, int lch, long long col, int lt, int npl) {
  constexpr int NW = NTH / 32;
  constexpr int NTA = (NSP_LSUB / 8 * NT + NW - 1) / NW;       // update tiles per warp and product
  constexpr int nt = NT, kp = 8 * NT, ks = kp + 4;
  const int k = P.k, nz = P.nz;
  const NspSm<NT> S(k, lch);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const bool tma = (k & 1) == 0;
  unsigned mph = (unsigned)S.par[1];
  const bool per_level = P.radius_v > 0.0;
  const double sW = sqrt((double)(k - 1));
  const NspLane L = nsp_lane(lane);
  double* Xg = P.X + col * nz * k;
  double* wa = S.wa;
  NSP_T0();
  const int lev_b = per_level ? lt : 0, lev_e = per_level ? lt + 1 : nz;
  double* Xt = S.Yp;                                               // [lch][ks], then the analysed [lch][k]
  auto request_x = [&](int l0, int nl) {
    if (tma) {
      fence_proxy_async();                                   // the iteration's generic stores to Y and T first
      if (tid == 0) mbar_expect_tx(S.mbar_x, (unsigned)(nl * k * 8));
      for (int l = tid; l < nl; l += NTH)
        bulk_g2s(S.ys + (unsigned)(l * ks * 8), Xg + (long long)(l0 + l) * k, (unsigned)(k * 8), S.mbar_x);
    } else {
      for (int l = warp; l < nl; l += NW)
        for (int j = lane; j < k; j += 32) Xt[l * ks + j] = Xg[(long long)(l0 + l) * k + j];
    }
    if (kp > k) for (int e = tid; e < nl * (kp - k); e += NTH) Xt[(e / (kp - k)) * ks + k + e % (kp - k)] = 0.0;
  };
  // v_out = Z v_in: warp per tile row, lanes read Z in the fragment pattern (row g, columns t, 4 + t of every tile) and
  // reduce over t.  (A version with every warp on independent loads -- tile COLUMNS per warp, partial row sums through
  // shared memory, < 1 k instead of 5.7 k cycles per product -- measured 1 % SLOWER end to end: the phase is hidden
  // behind the other CTA's products either way, and the wider version takes more issue slots and barriers away from
  // them.  A/B on one box, r02.)
  auto zmatvec = [&](const double* vin, double* vout) {
    for (int I = warp; I < nt; I += NW) {
      NspWalk zw;
      zw.start(I);
      double s = 0.0;
#pragma unroll 1
      for (int K = 0; K < nt; ++K) {
        const double z0 = lds_f64(zw.addr(S.zs, K, 0, L)), z1 = lds_f64(zw.addr(S.zs, K, 1, L));
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
  };
  // Ill-conditioned transforms (condition bound > NSP_REFINE_KAPPA): w = Z (Z g) inherits err(Z) sqrt(cond) -- g lies
  // along the LARGE eigenvalues of A, Z's error is relative to its largest entry 1 / sqrt(lambda_min) -- which is
  // what limited the packed kernel's condition range (2e-10 on the analysis mean at cond 7e4).  One step of iterative
  // refinement with Z Z as the approximate inverse squares that factor away: r = g - A w, w += Z (Z r), with
  // A w = shift w + Y^T (rho / sigma^2 o (Y w)) formed from the local rows themselves, re-gathered from L2 by bulk
  // copies NSP_RCH at a time into the (free) Y | T buffers.  Needs the selection of a single round (sel_row, sel_w).
  const bool refine = npl > 0 && S.sc[NSP_SC_KAPPA] > NSP_REFINE_KAPPA && S.sc[NSP_SC_ROUNDS] == 1.0;
  if (!refine) request_x(lev_b, min(lch, lev_e - lev_b));
  if (npl > 0) {
    zmatvec(S.gvec, S.tv);
    zmatvec(S.tv, wa);                                         // w = Z (Z g)
    if (refine) {
      unsigned gph = (unsigned)S.par[0];
      const double shift = (double)(k - 1) / P.inflation;
      double* Yr = S.Yp;                                         // [NSP_RCH][ks] re-gathered rows
      double* su = S.xm;                                         // [NSP_RCH] rho / sigma^2 * (Y w) of the round
      if (tid < k) S.tv[tid] = S.gvec[tid] - shift * wa[tid];    // r, accumulated in tv
      for (int c0 = 0; c0 < npl; c0 += NSP_RCH) {
        const int rows = min(NSP_RCH, npl - c0);
        if (tma) {
          fence_proxy_async();
          if (tid == 0) mbar_expect_tx(S.mbar_g, (unsigned)(rows * k * 8));
          for (int r = tid; r < rows; r += NTH)
            bulk_g2s(S.ys + (unsigned)(r * ks * 8), P.Yp + (long long)S.sel_row[c0 + r] * k, (unsigned)(k * 8), S.mbar_g);
          mbar_wait(S.mbar_g, gph);
          gph ^= 1u;
        } else {
          for (int r = warp; r < rows; r += NW) {
            const double* src = P.Yp + (long long)S.sel_row[c0 + r] * k;
            for (int j = lane; j < k; j += 32) Yr[r * ks + j] = src[j];
          }
        }
        __syncthreads();
        for (int r = warp; r < rows; r += NW) {                  // (Y w)_r, scaled by the row's weight
          double u = 0.0;
          for (int j = lane; j < k; j += 32) u = fma(Yr[r * ks + j], wa[j], u);
          u = warp_sum(u);
          if (lane == 0) su[r] = S.sel_w[c0 + r] * u;
        }
        __syncthreads();
        if (tid < k) {                                           // r -= Y^T su
          double acc = 0.0;
          for (int r = 0; r < rows; ++r) acc = fma(Yr[r * ks + tid], su[r], acc);
          S.tv[tid] -= acc;
        }
        __syncthreads();
      }
      if (tid == 0) S.par[0] = (int)gph;
      zmatvec(S.tv, S.gvec);                                     // (g is not needed any more)
      zmatvec(S.gvec, S.tv);
      if (tid < k) wa[tid] += S.tv[tid];
      __syncthreads();
      request_x(lev_b, min(lch, lev_e - lev_b));                 // (the rows above used the state block's staging)
    }
  }
  NSP_TICK(13);

  if (P.W_out && P.w_col == col && lt == 0) {
    for (int e = tid; e < k * k; e += NTH) {
      const int j = e / k, i = e - j * k;
      double vv;
      if (npl == 0) vv = (i == j) ? sqrt(P.inflation) : 0.0;
      else vv = wa[j] + sW * S.Zp[nsp_elem(j, i, nt)];
      P.W_out[e] = vv;
    }
  }

  for (int l0 = lev_b; l0 < lev_e; l0 += lch) {
    const int nl = min(lch, lev_e - l0);
    if (l0 > lev_b) {                                      // (more levels than one round stages)
      if (tid == 0) bulk_wait_read();
      __syncthreads();
      request_x(l0, nl);
    }
    if (tma) { mbar_wait(S.mbar_x, mph); mph ^= 1u; }
    __syncthreads();
    NSP_TICK2(11);
    // level means, perturbations in place, xbar + X' w: four lanes per level (members q, q + 4, ...), all levels
    // of the round at once
    for (int lb = 0; lb < nl; lb += NTH >> 2) {
      const int l = lb + (tid >> 2), q = tid & 3;
      const bool act = l < nl;                             // (whole warps stay in the shuffles)
      double* xr = Xt + (act ? l : 0) * ks;
      double s = 0.0;
      if (act) for (int j = q; j < k; j += 4) s += xr[j];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s /= (double)k;
      double m = 0.0;
      if (act)
        for (int j = q; j < k; j += 4) {
          const double xp = xr[j] - s;
          xr[j] = xp;
          if (npl > 0) m = fma(xp, wa[j], m);
        }
      m += __shfl_xor_sync(0xffffffffu, m, 1);
      m += __shfl_xor_sync(0xffffffffu, m, 2);
      if (act && q == 0) { S.xm[l] = s; S.ml[l] = s + m; }
    }
    __syncthreads();
    NSP_TICK2(12);
    if (npl == 0) {
      // no local observation: mean kept, perturbations inflated (LETKF.hpp:167-190); straight to HBM
      const double f = sqrt(P.inflation);
      for (int l = warp; l < nl; l += NW) {
        double s = 0.0;
        for (int j = lane; j < k; j += 32) {
          const double v = S.xm[l] + Xt[l * ks + j] * f;
          Xg[(long long)(l0 + l) * k + j] = v;
          s += v;
        }
        s = warp_sum(s);
        if (P.mean_out && lane == 0) P.mean_out[col * nz + l0 + l] = s * (1.0 / (double)k);
      }
      __syncthreads();                                     // (staging is refilled next)
      continue;
    }
    for (int s0 = 0; s0 < nl; s0 += NSP_LSUB) {
      const int ns = min(NSP_LSUB, nl - s0);
      const int ntr = (ns + 7) >> 3;
      const NsTiles<NTA> at = nsp_rect_tiles<NTA, NTH>(ntr, nt, warp);
      double uacc[NTA][2];
#pragma unroll
      for (int n = 0; n < NTA; ++n) { uacc[n][0] = 0.0; uacc[n][1] = 0.0; }
      if constexpr ((4 * NT) % NW == 0 && NT % NTA == 0 && NT <= NSP_FIXED_MAX_NT) {
        if (ntr == 4) {
          // full round: the warp's NTA tiles are row tile warp / (NT / NTA), column group warp % (NT / NTA)
          constexpr int NG = NT / NTA;
          const unsigned xa_s = S.ys + (unsigned)(((s0 + (warp / NG) * 8 + g) * ks + t) * 8);
          nsp_upd_dispatch<NT, NTA>(std::make_integer_sequence<int, NG>{}, warp % NG, xa_s, S.zs, L, uacc);
          goto update_product_done;
        }
      }
      {
      const double* xa = Xt + (s0 + g) * ks + t;
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
            b[n] = lds_f64(zw[n].addr(S.zs, K, h, L));
          }
#pragma unroll
          for (int n = 0; n < NTA; ++n) NSP_DMMA(uacc[n], a[n], b[n]);
        }
#pragma unroll
        for (int n = 0; n < NTA; ++n) zw[n].next(K, nt);
      }
      }
    update_product_done:
      __syncthreads();                                     // every warp is done with these rows of X'
      // the analysed rows go, dense, over the perturbations just consumed (row l at l k <= l ks)
#pragma unroll
      for (int n = 0; n < NTA; ++n)
        if (n < at.n) {
          const int l = s0 + at.ti[n] * 8 + g, i = at.tj[n] * 8 + 2 * t;
          if (l < s0 + ns) {
            if (i < k) Xt[l * k + i] = S.ml[l] + sW * uacc[n][0];
            if (i + 1 < k) Xt[l * k + i + 1] = S.ml[l] + sW * uacc[n][1];
          }
        }
    }
    NSP_TICK2(10);
    if (tma) fence_proxy_async();                          // generic writes -> visible to the bulk store
    __syncthreads();
    if (tma) {
      if (tid == 0) bulk_s2g(Xg + (long long)l0 * k, S.ys, (unsigned)(nl * k * 8));
    } else {
      for (int e = tid; e < nl * k; e += NTH) Xg[(long long)l0 * k + e] = Xt[e];
    }
    if (P.mean_out) {
      for (int lb = 0; lb < nl; lb += NTH >> 2) {
        const int l = lb + (tid >> 2), q = tid & 3;
        double s = 0.0;
        if (l < nl) for (int j = q; j < k; j += 4) s += Xt[l * k + j];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (l < nl && q == 0) P.mean_out[col * nz + l0 + l] = s * (1.0 / (double)k);
      }
    }
  }
  if (tid == 0) S.par[1] = (int)mph;     // (published by the next phase's barriers)
  NSP_TICK(14);
}

// WORK = true: consume the classifying pass's work list (per-level analyses).
template <int NT, int NTH, int MINB, bool WORK, bool EXT = false>
__global__ void __launch_bounds__(NTH, MINB) letkf_nsp_kernel(const __grid_constant__ ColParams P, int lch) {
  const int k = P.k, nz = P.nz;
  const int tid = threadIdx.x;
  {
    const NspSm<NT> S(k, lch);
    if (tid == 0) { mbar_init(S.mbar_g, 1); mbar_init(S.mbar_x, 1); S.par[0] = 0; S.par[1] = 0; }
  }
  __syncthreads();
  const int nxf = P.radius_v > 0.0 ? nz : 1;
  constexpr bool work = WORK;
  const long long ncols = work ? (long long)*P.work_count : (P.cols ? P.ncols : (long long)P.own_nx * P.own_ny);

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
    int col_iters = 0;
    long long col_npl = 0;
    bool col_fail = false;
    {
      // the column's state is first touched in phase 3: start pulling it towards L2 now
      const double* Xg = P.X + col * nz * k;
      for (int e = tid * 16; e < nz * k; e += NTH * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(Xg + e));
    }
    const int lt_e = work ? lt_b + 1 : nxf;
    for (int lt = lt_b; lt < lt_e; ++lt) {
#ifdef NSP_PROFILE
      if (threadIdx.x == 0) nsp_prof_ = P.stats;
      const long long nsp_col_t0 = clock64();
#endif
      const int r1 = nsp_phase_gram<NT, NTH, EXT>(P, lch, col, gx, gy, lt);
      const int npl = r1 & ((1 << 28) - 1), status = r1 >> 28;
      if (lt == lt_b) col_npl = npl;
      if (status != 0) continue;                               // handed to another kernel
      bool ok = true;
      if (npl > 0) {
        // ---------------- phase 2: Z = A^{-1/2}
        const NspSm<NT> S(k, lch);
        int rbuf = 0;
        const double shift = (double)(k - 1) / P.inflation;
        int it;
        if constexpr (NT <= NSP_FIXED_MAX_NT)
          it = nsp_inverse_sqrt<NT, NTH>(S.Zp, S.red, S.sc, rbuf, shift, rsqrt(shift), S.sc[NSP_SC_FRO], k);
        else
          it = nsp_inverse_sqrt_call<NT, NTH>(S.Zp, S.red, S.sc, rbuf, shift, rsqrt(shift), S.sc[NSP_SC_FRO], k);
        if (it == -2) {
          // condition bound beyond the limit: symmetric tiles are not trusted there, the full-product
          // kernel (k <= 80) or the Jacobi kernel redoes this transform
          if (tid == 0) {
            const unsigned slot = atomicAdd(P.redo_count, 1u);
            P.redo_items[slot] = col * nxf + lt;
          }
          continue;
        }
        if (it < 0) ok = false;
        col_iters = max(col_iters, it);
      }
      if (!ok) {
        col_fail = true;
        if (P.W_out && P.w_col == col && lt == 0)
          for (int e = tid; e < k * k; e += NTH) P.W_out[e] = nan("");
      } else {
        nsp_phase_update<NT, NTH>(P, lch, col, lt, npl);
      }
#if defined(NSP_PROFILE) && !defined(NSP_PROFILE_WARPS)
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
  if (tid == 0) bulk_wait_all();         // shared memory must outlive the last bulk store's reads
}

// levels staged per update round: all of them up to NSP_LCH_CAP (multiple of 8)
static int nsp_level_chunk(int k, int nz) {
  (void)k;
  return std::max(8, std::min(NSP_LCH_CAP, (nz + 7) & ~7));
}
static size_t nsp_smem_bytes(int k, int lch, int nth) {
  (void)nth;
  const size_t dbl = (size_t)nsp_region_doubles(k, lch) + 3 * (size_t)ns_kp(k) + 2 * NSP_LCH_CAP + 32 + NSP_SC_DOUBLES + 2 * NS_SELCAP + 2;
  return dbl * 8 + (size_t)NS_SELCAP * 4 + (64 + 4) * 4 + 16;
}
