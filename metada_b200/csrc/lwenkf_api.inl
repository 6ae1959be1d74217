// C-ABI entry point of the locally weighted EnKF (LWEnKF.hpp:207-334); included by mdc_api.cu.
// The dense P x P solve and the eigenvalues of S come from cuSOLVER (bound at run time like NCCL: the library loads
// without it and this entry point then reports MDC_ERR_UNSUPPORTED); everything else is this repo's kernels.
#include <dlfcn.h>
namespace {

struct CusolverApi {
  void* h = nullptr;
  int (*Create)(void**) = nullptr;
  int (*Destroy)(void*) = nullptr;
  int (*SetStream)(void*, cudaStream_t) = nullptr;
  int (*Dgetrf_bufferSize)(void*, int, int, double*, int, int*) = nullptr;
  int (*Dgetrf)(void*, int, int, double*, int, double*, int*, int*) = nullptr;
  int (*Dgetrs)(void*, int, int, int, const double*, int, const int*, double*, int, int*) = nullptr;
  int (*Dsyevd_bufferSize)(void*, int, int, int, const double*, int, const double*, int*) = nullptr;
  int (*Dsyevd)(void*, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
  bool load() {
    if (h) return true;
    for (const char* name : {"libcusolver.so.11", "libcusolver.so"}) { h = dlopen(name, RTLD_NOW); if (h) break; }
    if (!h) return false;
#define MDC_CS(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(h, sym)); if (!field) { h = nullptr; return false; }
    MDC_CS(Create, "cusolverDnCreate") MDC_CS(Destroy, "cusolverDnDestroy") MDC_CS(SetStream, "cusolverDnSetStream")
    MDC_CS(Dgetrf_bufferSize, "cusolverDnDgetrf_bufferSize") MDC_CS(Dgetrf, "cusolverDnDgetrf") MDC_CS(Dgetrs, "cusolverDnDgetrs")
    MDC_CS(Dsyevd_bufferSize, "cusolverDnDsyevd_bufferSize") MDC_CS(Dsyevd, "cusolverDnDsyevd")
#undef MDC_CS
    return true;
  }
};
CusolverApi g_cusolver;

// per-member sum of squared deviations from the stored mean -> host [k]
int member_sqnorms(mdc_ctx* ctx, mdc_ens* e, std::vector<double>& out) {
  const int k = e->k;
  const int64_t npts = (int64_t)e->nx * e->ny * e->nz;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((npts + 7) / 8, (int64_t)ctx->sm_count * 8));
  double *partial = nullptr, *sums = nullptr;
  if (tmp_alloc(ctx, &partial, (size_t)nb * k) || tmp_alloc(ctx, &sums, (size_t)k)) return MDC_ERR_CUDA;
  lw_member_sqnorm_kernel<<<nb, 256, 0, ctx->stream>>>(e->X, e->mean, npts, k, partial);
  MDC_LAUNCH_CHECK(ctx);
  reduce_partials_kernel<<<mdc_div_up(k, 128), 128, 0, ctx->stream>>>(partial, nb, k, sums);
  MDC_LAUNCH_CHECK(ctx);
  out.assign((size_t)k, 0.0);
  MDC_CUDA(ctx, cudaMemcpyAsync(out.data(), sums, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  tmp_free(ctx, partial); tmp_free(ctx, sums);
  return MDC_OK;
}

}  // namespace

extern "C" int mdc_lwenkf_analyse(mdc_ens* e, mdc_obs* o, double inflation, double loc_radius, int loc_fn, int weighting,
                                  const double* Z, uint64_t seed, mdc_lwenkf_diag* diag) {
  mdc_ctx* ctx = e->ctx;
  if (int rc = check_global_args(e, o, "lwenkf")) return rc;
  if (!(inflation > 0.0) || !(loc_radius > 0.0)) MDC_FAIL(ctx, MDC_ERR_INVALID, "lwenkf: inflation and localization_radius must be > 0");
  if (weighting < 0 || weighting > 3) MDC_FAIL(ctx, MDC_ERR_INVALID, "lwenkf: weighting scheme %d", weighting);
  if (loc_fn != MDC_LOC_CUTOFF && loc_fn != MDC_LOC_GAUSSIAN && loc_fn != MDC_LOC_EXPONENTIAL && loc_fn != MDC_LOC_REF_GASPARI_COHN)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "lwenkf: localisation function %d (cutoff, gaussian, exponential, ref_gaspari_cohn)", loc_fn);
  if (o->P > 46000) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "lwenkf: S is dense P x P; P = %lld is too many", (long long)o->P);
  if (!g_cusolver.load()) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "lwenkf: libcusolver not found (dense LU / eigenvalues of S)");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int k = e->k, ks = k | 1;
  const int64_t P = o->P, npts = (int64_t)e->nx * e->ny * e->nz, G = (int64_t)e->nx * e->ny;
  cudaStream_t s = ctx->stream;
  if (int rc = ens_mean_device(e)) return rc;                    // LWEnKF.hpp:219
  if (int rc = mdc_hx_idw4(e, o)) return rc;                     // :241-255
  // ---- weights (:400-436) on the host from per-member reductions
  std::vector<double> sq;
  if (int rc = member_sqnorms(ctx, e, sq)) return rc;
  std::vector<double> w((size_t)k);
  if (weighting == 3) {
    double* dq = nullptr;
    if (tmp_alloc(ctx, &dq, (size_t)k)) return MDC_ERR_CUDA;
    lw_likelihood_kernel<<<k, 256, 0, s>>>(o->Y, o->val, o->err, o->valid, P, k, dq);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaMemcpyAsync(w.data(), dq, (size_t)k * 8, cudaMemcpyDeviceToHost, s));
    MDC_CUDA(ctx, cudaStreamSynchronize(s));
    tmp_free(ctx, dq);
    for (int m = 0; m < k; ++m) w[m] = std::exp(-0.5 * w[m]);
  } else {
    for (int m = 0; m < k; ++m) w[m] = weighting == 0 ? 1.0 / k : weighting == 1 ? 1.0 / (std::sqrt(sq[m]) + 1e-8) : 1.0 / (sq[m] + 1e-8);
    if (weighting != 0) { double t = 0.0; for (double v : w) t += v; for (double& v : w) v /= t; }
  }
  { double t = 0.0; for (double v : w) t += v; for (double& v : w) v /= t; }
  double wmax = w[0], wmin = w[0], wvar = 0.0, bs = 0.0;
  for (int m = 0; m < k; ++m) { wmax = std::max(wmax, w[m]); wmin = std::min(wmin, w[m]); wvar += (w[m] - 1.0 / k) * (w[m] - 1.0 / k); bs += sq[m]; }
  wvar /= k;
  bool wbad = false;
  for (double v : w) if (!(v == v)) wbad = true;
  if (wbad) MDC_FAIL(ctx, MDC_ERR_NUMERIC, "lwenkf: the member weights are not finite (likelihood weights underflow to 0 / 0, as in the reference)");
  // ---- S, its LU, Gs = S^-1 Y', cond(S)
  double *dS = nullptr, *dS2 = nullptr, *dB = nullptr, *dGs = nullptr, *dD = nullptr, *dZ = nullptr, *dw = nullptr, *dsc = nullptr, *dev = nullptr,
         *dwork = nullptr, *dmm = nullptr;
  int *dipiv = nullptr, *dinfo = nullptr;
  void* cs = nullptr;
  auto cleanup = [&]() {
    tmp_free(ctx, dS); tmp_free(ctx, dS2); tmp_free(ctx, dB); tmp_free(ctx, dGs); tmp_free(ctx, dD); tmp_free(ctx, dZ); tmp_free(ctx, dw);
    tmp_free(ctx, dsc); tmp_free(ctx, dev); tmp_free(ctx, dwork); tmp_free(ctx, dmm); tmp_free(ctx, dipiv); tmp_free(ctx, dinfo);
    if (cs) g_cusolver.Destroy(cs);
  };
  if (tmp_alloc(ctx, &dS, (size_t)P * P) || tmp_alloc(ctx, &dS2, (size_t)P * P) || tmp_alloc(ctx, &dB, (size_t)P * k) ||
      tmp_alloc(ctx, &dGs, (size_t)P * k) || tmp_alloc(ctx, &dD, (size_t)P * k) || tmp_alloc(ctx, &dw, (size_t)k) ||
      tmp_alloc(ctx, &dsc, (size_t)16) || tmp_alloc(ctx, &dev, (size_t)P) || tmp_alloc(ctx, &dipiv, (size_t)P) || tmp_alloc(ctx, &dinfo, (size_t)2)) {
    cleanup();
    return MDC_ERR_CUDA;
  }
  cudaMemcpyAsync(dw, w.data(), (size_t)k * 8, cudaMemcpyHostToDevice, s);
  lw_build_S_kernel<<<grid_for(ctx, P * P, 256, 8), 256, (size_t)k * 8, s>>>(o->Yp, dw, o->err, o->valid, P, k, loc_fn, loc_radius, dS);
  ctx->launches++;
  cudaMemcpyAsync(dS2, dS, (size_t)P * P * 8, cudaMemcpyDeviceToDevice, s);
  lw_transpose_kernel<<<grid_for(ctx, P * k, 256, 8), 256, 0, s>>>(o->Yp, P, k, dB, 1);
  ctx->launches++;
  int rc = MDC_OK, lwork = 0, lwork2 = 0, info[2] = {0, 0};
  if (g_cusolver.Create(&cs) || g_cusolver.SetStream(cs, s)) { cleanup(); MDC_FAIL(ctx, MDC_ERR_CUDA, "lwenkf: cusolverDnCreate failed"); }
  // (S is symmetric up to rounding: its row-major and column-major images coincide)
  if (g_cusolver.Dgetrf_bufferSize(cs, (int)P, (int)P, dS, (int)P, &lwork) ||
      g_cusolver.Dsyevd_bufferSize(cs, 0 /*NOVECTOR*/, 0 /*LOWER*/, (int)P, dS2, (int)P, dev, &lwork2)) { cleanup(); MDC_FAIL(ctx, MDC_ERR_CUDA, "lwenkf: cuSOLVER workspace query failed"); }
  if (tmp_alloc(ctx, &dwork, (size_t)std::max(lwork, lwork2))) { cleanup(); return MDC_ERR_CUDA; }
  int st = g_cusolver.Dsyevd(cs, 0, 0, (int)P, dS2, (int)P, dev, dwork, lwork2, dinfo);          // eigenvalues: cond(S), :289-292
  if (!st) st = g_cusolver.Dgetrf(cs, (int)P, (int)P, dS, (int)P, dwork, dipiv, dinfo + 1);       // S.inverse() (:279) as an LU solve
  if (!st) st = g_cusolver.Dgetrs(cs, 0 /*N*/, (int)P, k, dS, (int)P, dipiv, dB, (int)P, dinfo + 1);
  cudaMemcpyAsync(info, dinfo, sizeof(info), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  if (st || info[1] != 0 || cudaGetLastError() != cudaSuccess) { cleanup(); MDC_FAIL(ctx, MDC_ERR_NUMERIC, "lwenkf: LU solve with S failed (status %d, info %d)", st, info[1]); }
  lw_transpose_kernel<<<grid_for(ctx, P * k, 256, 8), 256, 0, s>>>(dB, P, k, dGs, 0);
  ctx->launches++;
  std::vector<double> ev((size_t)P);
  cudaMemcpyAsync(ev.data(), dev, (size_t)P * 8, cudaMemcpyDeviceToHost, s);
  // ---- D = yo + sqrt(R) Z - Yb (:295-306), innovation norm
  if (Z) {
    if (tmp_alloc(ctx, &dZ, (size_t)P * k)) { cleanup(); return MDC_ERR_CUDA; }
    cudaMemcpyAsync(dZ, Z, (size_t)P * k * 8, cudaMemcpyHostToDevice, s);
  }
  enkf_innov_kernel<<<grid_for(ctx, P * k, 256, 8), 256, 0, s>>>(o->Y, dZ, o->val, o->err, P, k, seed, dD);
  ctx->launches++;
  obs_scalar_stats_kernel<<<1, 1024, 0, s>>>(o->d, o->err, o->valid, P, dsc);
  ctx->launches++;
  // ---- x_a = xb + x' sqrt(infl) + (K o Lg) D, K streamed
  const size_t smem = ((size_t)LW_TP * ks + 2 * (size_t)LW_TO * ks + (size_t)LW_TP * (LW_TO + 1) + LW_TP) * sizeof(double);
  cudaFuncSetAttribute(lw_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((npts + LW_TP - 1) / LW_TP, (int64_t)ctx->sm_count * 4));
  if (tmp_alloc(ctx, &dmm, (size_t)grid * 2)) { cleanup(); return MDC_ERR_CUDA; }
  lw_apply_kernel<<<grid, 256, smem, s>>>(e->X, e->mean, dGs, dD, npts, P, k, e->nz, G, std::sqrt(inflation), loc_fn, loc_radius, dmm);
  ctx->launches++;
  std::vector<double> mm((size_t)grid * 2);
  double hsc[3] = {0, 0, 0};
  cudaMemcpyAsync(mm.data(), dmm, mm.size() * 8, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(hsc, dsc, sizeof(hsc), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  if (cudaGetLastError() != cudaSuccess) { cleanup(); MDC_FAIL(ctx, MDC_ERR_CUDA, "lwenkf: update kernel failed"); }
  // ---- analysis statistics (:316-333)
  std::vector<double> sqa;
  rc = ens_mean_device(e);
  if (!rc) rc = member_sqnorms(ctx, e, sqa);
  if (!rc && diag) {
    double kmax = -INFINITY, kmin = INFINITY, lo = INFINITY, hi = 0.0, as = 0.0;
    for (int b = 0; b < grid; ++b) { kmax = std::max(kmax, mm[2 * b]); kmin = std::min(kmin, mm[2 * b + 1]); }
    for (double v : ev) { lo = std::min(lo, std::fabs(v)); hi = std::max(hi, std::fabs(v)); }
    for (double v : sqa) as += v;
    diag->innovation_norm = std::sqrt(hsc[0]);
    diag->background_spread = std::sqrt(inflation * bs / ((double)npts * k));
    diag->analysis_spread = std::sqrt(as / ((double)k * npts));
    diag->max_kalman_gain = kmax; diag->min_kalman_gain = kmin; diag->condition_number = hi / lo;
    diag->max_weight = wmax; diag->min_weight = wmin; diag->weight_variance = wvar;
  }
  cleanup();
  o->have_hx = false;
  return rc;
}
