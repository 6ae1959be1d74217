// C-ABI entry points of the global (non-localised) ETKF and EnKF; included by mdc_api.cu.
namespace {

int gram_reduce(mdc_ctx* ctx, mdc_obs* o, const double* B, int kb, double* d_out /*k*k + k*kb*/) {
  const int k = o->k;
  const int nout = k * k + k * kb;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((o->P + 255) / 256, ctx->sm_count * 2));
  double* partial = nullptr;
  if (tmp_alloc(ctx, &partial, (size_t)nb * nout)) return MDC_ERR_CUDA;
  size_t smem = ((size_t)2 * GK_ROWS * k + (size_t)GK_ROWS * kb) * sizeof(double);
  MDC_CUDA(ctx, cudaFuncSetAttribute(obs_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  obs_gram_kernel<<<nb, GK_THREADS, smem, ctx->stream>>>(o->Yp, B, kb, o->err, o->valid, o->P, k, partial);
  MDC_LAUNCH_CHECK(ctx);
  reduce_partials_kernel<<<mdc_div_up(nout, 128), 128, 0, ctx->stream>>>(partial, nb, nout, d_out);
  MDC_LAUNCH_CHECK(ctx);
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  tmp_free(ctx, partial);
  return MDC_OK;
}

int apply_transform(mdc_ctx* ctx, mdc_ens* e, const double* dW, double* spread2 /*host [2] or null*/) {
  const int k = e->k, ks = k | 1;
  const int64_t npts = (int64_t)e->nx * e->ny * e->nz;
  size_t smem = ((size_t)k * ks + 2 * (size_t)GA_TP * ks + GA_TP) * sizeof(double);
  if ((int)smem > ctx->max_smem_optin) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "global apply: k=%d needs %zu B shared memory", k, smem);
  MDC_CUDA(ctx, cudaFuncSetAttribute(global_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  MDC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, global_apply_kernel, GK_THREADS, smem));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((npts + GA_TP - 1) / GA_TP, (int64_t)ctx->sm_count * std::max(occ, 1)));
  double* dpart = nullptr;
  if (spread2 && tmp_alloc(ctx, &dpart, (size_t)grid * 2)) return MDC_ERR_CUDA;
  global_apply_kernel<<<grid, GK_THREADS, smem, ctx->stream>>>(e->X, e->mean, dW, npts, k, dpart);
  MDC_LAUNCH_CHECK(ctx);
  if (spread2) {
    std::vector<double> h((size_t)grid * 2);
    MDC_CUDA(ctx, cudaMemcpyAsync(h.data(), dpart, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    spread2[0] = spread2[1] = 0.0;
    for (int b = 0; b < grid; ++b) { spread2[0] += h[2 * b]; spread2[1] += h[2 * b + 1]; }
    tmp_free(ctx, dpart);
  }
  return MDC_OK;
}

int check_global_args(mdc_ens* e, mdc_obs* o, const char* who) {
  mdc_ctx* ctx = e->ctx;
  if (o->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "%s: ens/obs belong to different contexts", who);
  if (e->own_nx != e->nx || e->own_ny != e->ny || e->gnx != e->nx || e->gny != e->ny)
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "%s: the global (non-localised) filters do not shard; use one GPU", who);
  if (o->P <= 0) MDC_FAIL(ctx, MDC_ERR_INVALID, "%s: no observations", who);
  if (e->k > 128 || e->k < 2) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "%s: needs 2 <= k <= 128 members", who);
  return MDC_OK;
}

}  // namespace

extern "C" {

int mdc_etkf_analyse(mdc_ens* e, mdc_obs* o, double inflation) {
  mdc_ctx* ctx = e->ctx;
  if (int rc = check_global_args(e, o, "etkf")) return rc;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int k = e->k, ks = k | 1;
  if (int rc = ens_mean_device(e)) return rc;                    // ETKF.hpp:111
  if (int rc = mdc_hx_idw4(e, o)) return rc;                     // :128-141
  double *dG = nullptr, *dW = nullptr;
  if (tmp_alloc(ctx, &dG, (size_t)k * k + k) || tmp_alloc(ctx, &dW, (size_t)k * k)) return MDC_ERR_CUDA;
  int rc = gram_reduce(ctx, o, o->d, 1, dG);                      // C and g (:150, :155)
  if (!rc) {
    MDC_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    size_t smem = ((size_t)2 * k * ks + (size_t)k * k + k) * sizeof(double);
    MDC_CUDA(ctx, cudaFuncSetAttribute(global_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    global_solve_kernel<<<1, GK_THREADS, smem, ctx->stream>>>(dG, k, 0, inflation, dW, nullptr, ctx->d_flags);
    MDC_LAUNCH_CHECK(ctx);
    int flag = 0;
    MDC_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) { snprintf(ctx->err, sizeof(ctx->err), "etkf: ensemble-space matrix not SPD"); rc = MDC_ERR_NUMERIC; }
  }
  if (!rc) rc = apply_transform(ctx, e, dW, nullptr);             // :163-176
  if (!rc) { MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
  tmp_free(ctx, dG); tmp_free(ctx, dW);
  o->have_hx = false;
  return rc;
}

int mdc_enkf_analyse(mdc_ens* e, mdc_obs* o, double inflation, const double* Z, uint64_t seed,
                     int want_gain_stats, mdc_enkf_diag* diag) {
  mdc_ctx* ctx = e->ctx;
  if (int rc = check_global_args(e, o, "enkf")) return rc;
  if (!(inflation > 0.0)) MDC_FAIL(ctx, MDC_ERR_INVALID, "enkf: inflation must be > 0");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int k = e->k, ks = k | 1;
  const int64_t P = o->P, npts = (int64_t)e->nx * e->ny * e->nz;
  cudaStream_t s = ctx->stream;
  if (int rc = ens_mean_device(e)) return rc;                    // EnKF.hpp:149
  if (int rc = mdc_hx_idw4(e, o)) return rc;                     // :168-181
  double *dZ = nullptr, *dD = nullptr, *dG = nullptr, *dW = nullptr, *dAinv = nullptr, *dsc = nullptr;
  int rc = MDC_OK;
  auto cleanup = [&]() { tmp_free(ctx, dZ); tmp_free(ctx, dD); tmp_free(ctx, dG); tmp_free(ctx, dW); tmp_free(ctx, dAinv); tmp_free(ctx, dsc); };
  if (tmp_alloc(ctx, &dD, (size_t)P * k) || tmp_alloc(ctx, &dG, (size_t)2 * k * k) || tmp_alloc(ctx, &dW, (size_t)k * k) ||
      tmp_alloc(ctx, &dAinv, (size_t)k * k) || tmp_alloc(ctx, &dsc, (size_t)16 + k)) { cleanup(); return MDC_ERR_CUDA; }
  if (Z) {
    if (tmp_alloc(ctx, &dZ, (size_t)P * k)) { cleanup(); return MDC_ERR_CUDA; }
    cudaError_t ce = cudaMemcpyAsync(dZ, Z, (size_t)P * k * 8, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) { cleanup(); MDC_FAIL(ctx, MDC_ERR_CUDA, "enkf: upload of Z failed: %s", cudaGetErrorString(ce)); }
  }
  enkf_innov_kernel<<<grid_for(ctx, P * k, 256, 8), 256, 0, s>>>(o->Y, dZ, o->val, o->err, P, k, seed, dD);   // :212-222
  ctx->launches++;
  obs_scalar_stats_kernel<<<1, 1024, 0, s>>>(o->d, o->err, o->valid, P, dsc);                                 // :184
  ctx->launches++;
  rc = gram_reduce(ctx, o, dD, k, dG);                           // C = Y'^T R^-1 Y', F = Y'^T R^-1 D
  if (!rc) {
    cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s);
    size_t smem = ((size_t)2 * k * ks + (size_t)k * k + k) * sizeof(double);
    cudaFuncSetAttribute(global_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    global_solve_kernel<<<1, GK_THREADS, smem, s>>>(dG, k, 1, inflation, dW, dAinv, ctx->d_flags);
    ctx->launches++;
    int flag = 0;
    cudaMemcpyAsync(&flag, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (cudaGetLastError() != cudaSuccess || flag) { snprintf(ctx->err, sizeof(ctx->err), "enkf: ensemble-space solve failed"); rc = MDC_ERR_NUMERIC; }
  }
  double kmax = NAN, kmin = NAN, cond = NAN;
  double hsc[3] = {0, 0, 0};
  if (!rc) {
    cudaMemcpyAsync(hsc, dsc, sizeof(hsc), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  }
  if (!rc && want_gain_stats) {
    // K = sqrt(infl) X' A^-1 Y'^T R^-1, streamed max/min (:199-203)
    double *dM = nullptr, *dmm = nullptr;
    if (tmp_alloc(ctx, &dM, (size_t)P * k) || tmp_alloc(ctx, &dmm, 2)) { cleanup(); return MDC_ERR_CUDA; }
    double init[2] = {-INFINITY, INFINITY};
    cudaMemcpyAsync(dmm, init, sizeof(init), cudaMemcpyHostToDevice, s);
    enkf_gain_factor_kernel<<<grid_for(ctx, P * k, 256, 8), 256, 0, s>>>(o->Yp, dAinv, o->err, o->valid, P, k, dM);
    ctx->launches++;
    const int gkc = std::min((k + 3) & ~3, GM_KC), gks = ((gkc + 7) & ~7) + 4;
    size_t smem = (size_t)(GM_TP + GM_TO) * gks * sizeof(double);
    cudaFuncSetAttribute(enkf_gain_minmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)((npts + GM_TP - 1) / GM_TP), (unsigned)((P + GM_TO - 1) / GM_TO));
    enkf_gain_minmax_kernel<<<grid, GK_THREADS, smem, s>>>(e->X, e->mean, dM, npts, P, k, std::sqrt(inflation), dmm);
    ctx->launches++;
    double hmm[2];
    cudaMemcpyAsync(hmm, dmm, sizeof(hmm), cudaMemcpyDeviceToHost, s);
    // cond(S) (:206-209): for R = sigma^2 I the spectrum of S = R + Y'Y'^T/(k-1) is
    // sigma^2 (1 + lambda_i(C)/(k-1)) on range(Y') and sigma^2 elsewhere, lambda_i(C) from a k x k
    // Jacobi eigensolve; for non-uniform R it is not available in ensemble space -> NaN.
    std::vector<double> ev((size_t)k);
    size_t sm2 = (size_t)k * ks * sizeof(double);
    cudaFuncSetAttribute(sym_eigvals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    sym_eigvals_kernel<<<1, GK_THREADS, sm2, s>>>(dG, k, dsc + 16);
    ctx->launches++;
    cudaMemcpyAsync(ev.data(), dsc + 16, (size_t)k * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (cudaGetLastError() != cudaSuccess) { tmp_free(ctx, dM); tmp_free(ctx, dmm); cleanup(); MDC_FAIL(ctx, MDC_ERR_CUDA, "enkf: gain statistics kernels failed"); }
    kmax = hmm[0]; kmin = hmm[1];
    if (hsc[1] == hsc[2] && hsc[1] > 0.0) {
      std::sort(ev.begin(), ev.end(), [](double a, double b) { return a > b; });
      const double km1 = (double)(k - 1);
      const double lmax = 1.0 + ev[0] / km1;
      const double lmin = (P >= k) ? 1.0 : 1.0 + ev[(size_t)P - 1] / km1;
      cond = lmax / lmin;
    }
    tmp_free(ctx, dM); tmp_free(ctx, dmm);
  }
  double sp[2] = {0, 0};
  if (!rc) rc = apply_transform(ctx, e, dW, sp);                 // :215-234
  if (!rc && diag) {
    diag->innovation_norm = std::sqrt(hsc[0]);
    diag->background_spread = std::sqrt(inflation * sp[0] / ((double)npts * k));   // :334
    diag->analysis_spread = std::sqrt(sp[1] / ((double)k * npts));                 // :246-253
    diag->max_kalman_gain = kmax;
    diag->min_kalman_gain = kmin;
    diag->condition_number = cond;
  }
  cleanup();
  o->have_hx = false;
  return rc;
}

}  // extern "C"
