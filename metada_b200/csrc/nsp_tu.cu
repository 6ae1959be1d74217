// One group of instantiations of letkf_nsp_kernel (see nsp_launch.h).  Build: nvcc -c -DNSP_LO=a -DNSP_HI=b ...
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "nsp_launch.h"

namespace {   // internal linkage for everything the kernel headers define (they are also part of mdc_api.cu)
#include "letkf_nsp.cuh"

template <typename K>
int launchp(K kern, int nth, const ColParams& cp, int lch, int sms, long long total_cols, mdc_ctx* ctx) {
  const size_t smemp = nsp_smem_bytes(cp.k, lch, nth);
  if ((int)smemp > ctx->max_smem_optin)
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: k=%d needs %zu B shared memory > %d available", cp.k, smemp, ctx->max_smem_optin);
  MDC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp));
  int occ = 1;
  MDC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nth, smemp));
  if (occ < 1) occ = 1;
  if (const char* e = getenv("MDC_NSP_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));   // development: co-residency experiments
  int grid = (int)std::max<long long>(1, std::min<long long>(total_cols, (long long)sms * occ));
  kern<<<grid, nth, smemp, ctx->stream>>>(cp, lch);
  MDC_LAUNCH_CHECK(ctx);
  return MDC_OK;
}

#ifndef NSP_DEV_NTH10   /* development: threads per CTA for 41 <= k <= 80 */
#define NSP_DEV_NTH10 256
#endif
template <int NT>
int launch_nt(const ColParams& cp, int lch, int sms, int ext, int work, long long total_cols, mdc_ctx* ctx) {
  // k <= 80: 8 warps, two columns per SM (nt = 10 with 384 threads measured 6 % slower); above: 16 warps, one per SM
#ifdef NSP_DEV_MINB1      // development: one CTA per SM, no register cap (what ptxas does with the products then)
  constexpr int NTH = NT <= 10 ? 256 : 512, MINB = 1;
#else
  // k <= 40: 4 warps, four columns per SM (products of 15 tiles: the fixed cost per product dominates, more CTAs
  // interleave); k <= 80: 8 warps, two per SM (nt = 10 with 384 threads measured 6 % slower); above: 16 warps, one
  constexpr int NTH = NT <= 5 ? 128 : NT <= 10 ? NSP_DEV_NTH10 : 512, MINB = NT <= 5 ? 4 : NT <= 10 ? 2 : 1;
#endif
  if (ext && work) return launchp(letkf_nsp_kernel<NT, NTH, MINB, true, true>, NTH, cp, lch, sms, total_cols, ctx);
  if (ext) return launchp(letkf_nsp_kernel<NT, NTH, MINB, false, true>, NTH, cp, lch, sms, total_cols, ctx);
  if (work) return launchp(letkf_nsp_kernel<NT, NTH, MINB, true>, NTH, cp, lch, sms, total_cols, ctx);
  return launchp(letkf_nsp_kernel<NT, NTH, MINB, false>, NTH, cp, lch, sms, total_cols, ctx);
}

template <int NT>
int dispatch(int nt, const ColParams& cp, int lch, int sms, int ext, int work, long long total_cols, mdc_ctx* ctx) {
  if constexpr (NT > NSP_HI) return NSP_NOT_MINE;
  else {
    if (nt == NT) return launch_nt<NT>(cp, lch, sms, ext, work, total_cols, ctx);
    return dispatch<NT + 1>(nt, cp, lch, sms, ext, work, total_cols, ctx);
  }
}
}  // namespace

#define NSP_CAT2(a, b, c) nsp_launch_##a##_##b
#define NSP_CAT(a, b) NSP_CAT2(a, b, 0)
int NSP_CAT(NSP_LO, NSP_HI)(int nt, const void* colparams, int lch, int sms, int ext, int work, long long total_cols, mdc_ctx* ctx) {
  if (nt < NSP_LO || nt > NSP_HI) return NSP_NOT_MINE;
  return dispatch<NSP_LO>(nt, *static_cast<const ColParams*>(colparams), lch, sms, ext, work, total_cols, ctx);
}
