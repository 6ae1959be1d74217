// Launchers of the packed Newton-Schulz column kernel (letkf_nsp.cuh), one translation unit per group of tile
// counts (nsp_tu.cu compiled with -DNSP_LO / -DNSP_HI) so that the 42 instantiations build in parallel.
// `colparams` points at a ColParams (passed untyped: the kernel headers live in an unnamed namespace per unit).
// Return MDC_* codes, or NSP_NOT_MINE when nt = ceil(k / 8) is outside the unit's range.
#pragma once
#include "mdc_internal.cuh"

#define NSP_NOT_MINE (-1000)
#define NSP_GROUPS(X) X(3, 6) X(7, 9) X(10, 10) X(11, 12) X(13, 14) X(15, 16)

#define NSP_DECL(LO, HI) \
  int nsp_launch_##LO##_##HI(int nt, const void* colparams, int lch, int sms, int ext, int work, long long total_cols, mdc_ctx* ctx);
NSP_GROUPS(NSP_DECL)
#undef NSP_DECL
