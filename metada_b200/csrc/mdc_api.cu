// C ABI of the sm_100a METADA analysis backend (see include/metada_cuda_c_api.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared ...
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "bench_kernels.cuh"
#include "ens_kernels.cuh"
#include "geo_kernels.cuh"
#include "global_kernels.cuh"
#include "lwenkf_kernels.cuh"
#include "hx_kernels.cuh"
#include "index_kernels.cuh"
#include "letkf_kernels.cuh"
#include "letkf_ns.cuh"
#include "letkf_nsp.cuh"
#include "nsp_launch.h"
#include "letkf_v2.cuh"
#include "metrics_kernels.cuh"
#include "mdc_internal.cuh"

namespace {

constexpr size_t kStageBytes = 512ull << 20;  // staging for host<->device member transposes
constexpr int kMemberBatch = 8;

int grid_for(mdc_ctx* ctx, int64_t work_items, int threads, int per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

template <typename T>
int dev_alloc(mdc_ctx* ctx, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  MDC_CUDA(ctx, cudaMalloc((void**)p, n * sizeof(T)));
  return MDC_OK;
}

// Per-call temporaries of the global (ETKF / EnKF) paths: stream-ordered pool allocations.  Plain
// cudaMalloc / cudaFree map and unmap device memory on every call, which made a 3 ms EnKF analysis take
// 12 - 55 ms depending on what the process had allocated before (measured, tools/dbg_c2.py).
template <typename T>
int tmp_alloc(mdc_ctx* ctx, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  MDC_CUDA(ctx, cudaMallocAsync((void**)p, n * sizeof(T), ctx->stream));
  return MDC_OK;
}
template <typename T>
void tmp_free(mdc_ctx* ctx, T* p) {
  if (p) cudaFreeAsync((void*)p, ctx->stream);
}

__global__ void small_copy_kernel(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}
__global__ void fill4_kernel(int* dst, int a, int b, int c, int d) {
  if (threadIdx.x == 0) { dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; }
}

// Read `bytes` (a multiple of 8, <= 256) of device memory back to the host and synchronise the stream -- without the
// copy engine: inside the streamed pipeline its queue holds tens of milliseconds of member transfers.
int read_small(mdc_ctx* ctx, const void* dsrc, void* hdst, size_t bytes) {
  if (!ctx->h_small) {
    MDC_CUDA(ctx, cudaHostAlloc(&ctx->h_small, 256, cudaHostAllocMapped));
    MDC_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_small, ctx->h_small, 0));
  }
  small_copy_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long*)dsrc, (unsigned long long*)ctx->d_small, (int)(bytes / 8));
  MDC_LAUNCH_CHECK(ctx);
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(hdst, ctx->h_small, bytes);
  return MDC_OK;
}

int ensure_copy_stream(mdc_ctx* ctx) {
  if (ctx->copy_stream) return MDC_OK;
  MDC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    MDC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_full[i], cudaEventDisableTiming));
    MDC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
  }
  MDC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
  return MDC_OK;
}

int ensure_stage(mdc_ens* e, size_t elems) {
  if (e->stage_elems >= elems) return MDC_OK;
  if (e->stage) cudaFree(e->stage);
  e->stage = nullptr;
  e->stage_elems = 0;
  MDC_CUDA(e->ctx, cudaMalloc((void**)&e->stage, elems * sizeof(double)));
  e->stage_elems = elems;
  return MDC_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------ context
int mdc_ctx_create(int device, mdc_ctx** out) {
  if (!out) return MDC_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MDC_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= ndev) return MDC_ERR_INVALID;
  mdc_ctx* ctx = new (std::nothrow) mdc_ctx();
  if (!ctx) return MDC_ERR_INVALID;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MDC_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return MDC_ERR_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MDC_ERR_CUDA; }
  {   // keep freed temporaries in the device's default pool (tmp_alloc / tmp_free)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  for (auto& e : ctx->pe) cudaEventCreate(&e);
  if (cudaMalloc((void**)&ctx->d_flags, 16 * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_stats, 16 * sizeof(long long)) != cudaSuccess) {
    delete ctx;
    return MDC_ERR_CUDA;
  }
  cudaMemset(ctx->d_flags, 0, 16 * sizeof(int));
  cudaMemset(ctx->d_stats, 0, 16 * sizeof(long long));
  *out = ctx;
  return MDC_OK;
}

int mdc_ctx_destroy(mdc_ctx* ctx) {
  if (!ctx) return MDC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  cudaFree(ctx->d_flags);
  if (ctx->redo_items) cudaFree(ctx->redo_items);
  cudaFree(ctx->d_stats);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  for (auto& e : ctx->pe) cudaEventDestroy(e);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_full[i]); cudaEventDestroy(ctx->ev_free[i]); }
    cudaEventDestroy(ctx->ev_start);
    cudaStreamDestroy(ctx->copy_stream);
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return MDC_OK;
}

const char* mdc_last_error(const mdc_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int mdc_ctx_sync(mdc_ctx* ctx) {
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDC_OK;
}
void* mdc_ctx_stream(mdc_ctx* ctx) { return (void*)ctx->stream; }
int mdc_timer_start(mdc_ctx* ctx) {
  MDC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return MDC_OK;
}
int mdc_timer_stop(mdc_ctx* ctx, float* ms) {
  MDC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  MDC_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  MDC_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return MDC_OK;
}
int64_t mdc_ctx_launch_count(const mdc_ctx* ctx) { return ctx->launches; }
int mdc_ctx_sm_count(const mdc_ctx* ctx) { return ctx->sm_count; }
int mdc_ctx_last_stats(mdc_ctx* ctx, int64_t out[16]) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  long long h[16];
  if (int rc = read_small(ctx, ctx->d_stats, h, sizeof(h))) return rc;
  for (int i = 0; i < 16; ++i) out[i] = h[i];
  return MDC_OK;
}

int mdc_ctx_flush_l2(mdc_ctx* ctx) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->flush_buf) {
    ctx->flush_bytes = 256ull << 20;  // 2x the 126 MB L2
    MDC_CUDA(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  int64_t n = (int64_t)(ctx->flush_bytes / sizeof(float4));
  flush_l2_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((float4*)ctx->flush_buf, n);
  MDC_LAUNCH_CHECK(ctx);
  return MDC_OK;
}

int mdc_dev_malloc(mdc_ctx* ctx, int64_t bytes, void** out) {
  if (!out || bytes < 0) return MDC_ERR_INVALID;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  MDC_CUDA(ctx, cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 8)));
  return MDC_OK;
}
int mdc_dev_free(mdc_ctx* ctx, void* p) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  MDC_CUDA(ctx, cudaFree(p));
  return MDC_OK;
}

int mdc_dev_copy(mdc_ctx* ctx, void* dst, const void* src, int64_t bytes, int kind) {
  if (!dst || !src || bytes < 0 || (kind != 1 && kind != 2)) MDC_FAIL(ctx, MDC_ERR_INVALID, "dev_copy: bad arguments");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  MDC_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDC_OK;
}

// ------------------------------------------------------------------------------ ensemble
int mdc_ens_create(mdc_ctx* ctx, int nx, int ny, int nz, int k, mdc_ens** out) {
  if (!ctx || !out) return MDC_ERR_INVALID;
  *out = nullptr;
  if (nx <= 0 || ny <= 0 || nz <= 0 || k < 1) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_ens_create: bad dims %d %d %d k=%d", nx, ny, nz, k);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  mdc_ens* e = new (std::nothrow) mdc_ens();
  if (!e) MDC_FAIL(ctx, MDC_ERR_INVALID, "out of host memory");
  e->ctx = ctx;
  e->nx = nx; e->ny = ny; e->nz = nz; e->k = k; e->ny_cap = ny;
  e->gx0 = 0; e->gy0 = 0; e->gnx = nx; e->gny = ny; e->own_nx = nx; e->own_ny = ny;
  size_t total = (size_t)nx * ny * nz * k;
  cudaError_t ce = cudaMalloc((void**)&e->X, total * sizeof(double));
  if (ce != cudaSuccess) {
    delete e;
    MDC_FAIL(ctx, MDC_ERR_CUDA, "mdc_ens_create: cudaMalloc(%zu bytes) failed: %s", total * sizeof(double), cudaGetErrorString(ce));
  }
  *out = e;
  return MDC_OK;
}

int mdc_ens_destroy(mdc_ens* e) {
  if (!e) return MDC_OK;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(e->ctx->stream);
  cudaFree(e->X);
  if (e->mean) cudaFree(e->mean);
  if (e->stage) cudaFree(e->stage);
  cudaFree(e->glat); cudaFree(e->glon); cudaFree(e->vcoord); cudaFree(e->levmap);
  for (int l = 0; l < 2; ++l) { cudaFree(e->gc_start[l]); cudaFree(e->gc_pts[l]); cudaFree(e->gc_plat[l]); cudaFree(e->gc_plon[l]); }
  delete e;
  return MDC_OK;
}

int mdc_ens_set_domain(mdc_ens* e, int gx0, int gy0, int gnx, int gny, int own_nx, int own_ny) {
  mdc_ctx* ctx = e->ctx;
  if (gnx <= 0 || gny <= 0 || gx0 < 0 || gy0 < 0 || gx0 + e->nx > gnx || gy0 + e->ny > gny ||
      own_nx <= 0 || own_ny <= 0 || own_nx > e->nx || own_ny > e->ny)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_ens_set_domain: inconsistent domain");
  e->gx0 = gx0; e->gy0 = gy0; e->gnx = gnx; e->gny = gny; e->own_nx = own_nx; e->own_ny = own_ny;
  return MDC_OK;
}

int mdc_ens_set_rows(mdc_ens* e, int ny) {
  mdc_ctx* ctx = e->ctx;
  if (ny <= 0 || ny > e->ny_cap) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_rows: %d rows requested, %d allocated", ny, e->ny_cap);
  if (ny != e->ny) {
    e->ny = ny;
    e->gx0 = 0; e->gy0 = 0; e->gnx = e->nx; e->gny = ny; e->own_nx = e->nx; e->own_ny = ny;
  }
  return MDC_OK;
}

int mdc_ens_upload_members(mdc_ens* e, int m0, int count, const double* const* hosts) {
  mdc_ctx* ctx = e->ctx;
  if (m0 < 0 || count <= 0 || m0 + count > e->k) MDC_FAIL(ctx, MDC_ERR_INVALID, "upload: member range [%d,%d) out of range", m0, m0 + count);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t G = (int64_t)e->nx * e->ny, n = G * e->nz;
  for (int mb = 0; mb < count; mb += kMemberBatch) {
    const int cb = std::min(kMemberBatch, count - mb);
    const int64_t piece = std::min<int64_t>(n, (int64_t)(kStageBytes / sizeof(double)) / cb);
    if (int rc = ensure_stage(e, (size_t)piece * cb)) return rc;
    for (int64_t p0 = 0; p0 < n; p0 += piece) {
      const int64_t np = std::min(piece, n - p0);
      for (int c = 0; c < cb; ++c)
        MDC_CUDA(ctx, cudaMemcpyAsync(e->stage + (int64_t)c * np, hosts[mb + c] + p0, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      ens_scatter_members_kernel<<<mdc_div_up(np, 256), 256, 0, ctx->stream>>>(e->X, e->stage, p0, np, G, e->nz, e->k, m0 + mb, cb);
      MDC_LAUNCH_CHECK(ctx);
    }
  }
  return MDC_OK;
}

int mdc_ens_upload_member(mdc_ens* e, int m, const double* host) {
  const double* h[1] = {host};
  return mdc_ens_upload_members(e, m, 1, h);
}

int mdc_ens_download_members(mdc_ens* e, int m0, int count, double* const* hosts) {
  mdc_ctx* ctx = e->ctx;
  if (m0 < 0 || count <= 0 || m0 + count > e->k) MDC_FAIL(ctx, MDC_ERR_INVALID, "download: member range out of range");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t G = (int64_t)e->nx * e->ny, n = G * e->nz;
  for (int mb = 0; mb < count; mb += kMemberBatch) {
    const int cb = std::min(kMemberBatch, count - mb);
    const int64_t piece = std::min<int64_t>(n, (int64_t)(kStageBytes / sizeof(double)) / cb);
    if (int rc = ensure_stage(e, (size_t)piece * cb)) return rc;
    for (int64_t p0 = 0; p0 < n; p0 += piece) {
      const int64_t np = std::min(piece, n - p0);
      ens_gather_members_kernel<<<mdc_div_up(np, 256), 256, 0, ctx->stream>>>(e->X, e->stage, p0, np, G, e->nz, e->k, m0 + mb, cb);
      MDC_LAUNCH_CHECK(ctx);
      for (int c = 0; c < cb; ++c)
        MDC_CUDA(ctx, cudaMemcpyAsync(hosts[mb + c] + p0, e->stage + (int64_t)c * np, np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      // the staging buffer is reused by the next piece: same stream => ordered
    }
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDC_OK;
}

int mdc_ens_upload_members_rows(mdc_ens* e, int m0, int count, const double* const* hosts,
                                int host_ny, int host_y0) {
  mdc_ctx* ctx = e->ctx;
  if (m0 < 0 || count <= 0 || m0 + count > e->k) MDC_FAIL(ctx, MDC_ERR_INVALID, "upload_rows: member range out of range");
  if (host_y0 < 0 || host_y0 + e->ny > host_ny) MDC_FAIL(ctx, MDC_ERR_INVALID, "upload_rows: rows [%d,%d) outside the host array (%d rows)", host_y0, host_y0 + e->ny, host_ny);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t G = (int64_t)e->nx * e->ny, n = G * e->nz;
  const size_t width = (size_t)e->ny * e->nx * sizeof(double), spitch = (size_t)host_ny * e->nx * sizeof(double);
  // Double-buffered staging: the H2D copies of batch b + 1 run on the copy stream while the transpose of batch b
  // runs on the context's stream (one buffer and one stream serialised them: 6.2 ms per 8-member batch of a C5 slab
  // instead of the 4.1 ms the copy alone takes).
  if (int rc = ensure_copy_stream(ctx)) return rc;
  const size_t half = (size_t)n * std::min(kMemberBatch, count);
  if (int rc = ensure_stage(e, 2 * half)) return rc;
  cudaStream_t ms = ctx->stream, cs = ctx->copy_stream;
  MDC_CUDA(ctx, cudaEventRecord(ctx->ev_start, ms));            // earlier work on the stream may still read the staging buffer
  MDC_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_start, 0));
  int b = 0;
  for (int mb = 0; mb < count; mb += kMemberBatch, ++b) {
    const int cb = std::min(kMemberBatch, count - mb), h = b & 1;
    double* st = e->stage + (size_t)h * half;
    if (b >= 2) MDC_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_free[h], 0));   // the transpose of batch b - 2 has read this half
    for (int c = 0; c < cb; ++c)
      MDC_CUDA(ctx, cudaMemcpy2DAsync(st + (int64_t)c * n, width, hosts[mb + c] + (int64_t)host_y0 * e->nx, spitch, width,
                                      (size_t)e->nz, cudaMemcpyHostToDevice, cs));
    MDC_CUDA(ctx, cudaEventRecord(ctx->ev_full[h], cs));
    MDC_CUDA(ctx, cudaStreamWaitEvent(ms, ctx->ev_full[h], 0));
    ens_scatter_members_kernel<<<mdc_div_up(n, 256), 256, 0, ms>>>(e->X, st, 0, n, G, e->nz, e->k, m0 + mb, cb);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaEventRecord(ctx->ev_free[h], ms));
  }
  return MDC_OK;   // everything the caller may wait for is on the context's stream
}

int mdc_ens_download_members_rows(mdc_ens* e, int m0, int count, double* const* hosts, int host_ny,
                                  int host_y0, int nrows) {
  mdc_ctx* ctx = e->ctx;
  if (m0 < 0 || count <= 0 || m0 + count > e->k) MDC_FAIL(ctx, MDC_ERR_INVALID, "download_rows: member range out of range");
  if (nrows <= 0 || nrows > e->ny || host_y0 < 0 || host_y0 + nrows > host_ny) MDC_FAIL(ctx, MDC_ERR_INVALID, "download_rows: bad row range");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t G = (int64_t)e->nx * e->ny, n = G * e->nz;
  const size_t spitch = (size_t)e->ny * e->nx * sizeof(double), dpitch = (size_t)host_ny * e->nx * sizeof(double);
  const size_t width = (size_t)nrows * e->nx * sizeof(double);
  // double-buffered like the upload: the transpose of batch b + 1 overlaps the D2H copies of batch b
  if (int rc = ensure_copy_stream(ctx)) return rc;
  const size_t half = (size_t)n * std::min(kMemberBatch, count);
  if (int rc = ensure_stage(e, 2 * half)) return rc;
  cudaStream_t ms = ctx->stream, cs = ctx->copy_stream;
  int b = 0;
  for (int mb = 0; mb < count; mb += kMemberBatch, ++b) {
    const int cb = std::min(kMemberBatch, count - mb), h = b & 1;
    double* st = e->stage + (size_t)h * half;
    if (b >= 2) MDC_CUDA(ctx, cudaStreamWaitEvent(ms, ctx->ev_free[h], 0));   // the copies of batch b - 2 have drained this half
    ens_gather_members_kernel<<<mdc_div_up(n, 256), 256, 0, ms>>>(e->X, st, 0, n, G, e->nz, e->k, m0 + mb, cb);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaEventRecord(ctx->ev_full[h], ms));
    MDC_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_full[h], 0));
    for (int c = 0; c < cb; ++c)
      MDC_CUDA(ctx, cudaMemcpy2DAsync(hosts[mb + c] + (int64_t)host_y0 * e->nx, dpitch, st + (int64_t)c * n, spitch, width,
                                      (size_t)e->nz, cudaMemcpyDeviceToHost, cs));
    MDC_CUDA(ctx, cudaEventRecord(ctx->ev_free[h], cs));
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(cs));
  MDC_CUDA(ctx, cudaStreamSynchronize(ms));
  return MDC_OK;
}

int mdc_ens_download_member(mdc_ens* e, int m, double* host) {
  double* h[1] = {host};
  return mdc_ens_download_members(e, m, 1, h);
}

int mdc_ens_fill_synthetic(mdc_ens* e, uint64_t seed) {
  mdc_ctx* ctx = e->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t total = (int64_t)e->nx * e->ny * e->nz * e->k;
  ens_fill_synthetic_kernel<<<grid_for(ctx, total, 256, 16), 256, 0, ctx->stream>>>(
      e->X, e->nx, e->ny, e->nz, e->k, e->gx0, e->gy0, e->gnx, e->gny, seed);
  MDC_LAUNCH_CHECK(ctx);
  return MDC_OK;
}

static int ens_mean_device(mdc_ens* e) {
  mdc_ctx* ctx = e->ctx;
  const int64_t npts = (int64_t)e->nx * e->ny * e->nz;
  if (!e->mean) MDC_CUDA(ctx, cudaMalloc((void**)&e->mean, (size_t)e->nx * e->ny_cap * e->nz * sizeof(double)));
  constexpr int W = 4;
  size_t smem = (size_t)W * 32 * (e->k | 1) * sizeof(double);
  MDC_CUDA(ctx, cudaFuncSetAttribute(ens_mean_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((npts + W * 32 - 1) / (W * 32), (int64_t)ctx->sm_count * 4));
  ens_mean_kernel<W><<<grid, W * 32, smem, ctx->stream>>>(e->X, e->mean, npts, e->k);
  MDC_LAUNCH_CHECK(ctx);
  return MDC_OK;
}

int mdc_ens_mean(mdc_ens* e, double* host_mean) {
  mdc_ctx* ctx = e->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = ens_mean_device(e)) return rc;
  if (host_mean) {
    const int64_t G = (int64_t)e->nx * e->ny, npts = G * e->nz;
    if (int rc = ensure_stage(e, (size_t)npts)) return rc;
    mean_to_host_order_kernel<<<mdc_div_up(npts, 256), 256, 0, ctx->stream>>>(e->mean, e->stage, G, e->nz);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaMemcpyAsync(host_mean, e->stage, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MDC_OK;
}

int mdc_ens_checksum(mdc_ens* e, double* sum, double* sumsq) {
  mdc_ctx* ctx = e->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  double* d2 = nullptr;
  MDC_CUDA(ctx, cudaMalloc((void**)&d2, 2 * sizeof(double)));
  MDC_CUDA(ctx, cudaMemsetAsync(d2, 0, 2 * sizeof(double), ctx->stream));
  int64_t total = (int64_t)e->nx * e->ny * e->nz * e->k;
  ens_checksum_kernel<<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(e->X, total, d2);
  MDC_LAUNCH_CHECK(ctx);
  double h[2];
  MDC_CUDA(ctx, cudaMemcpyAsync(h, d2, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d2);
  if (sum) *sum = h[0];
  if (sumsq) *sumsq = h[1];
  return MDC_OK;
}

double* mdc_ens_devptr(mdc_ens* e) { return e->X; }
int64_t mdc_ens_bytes(const mdc_ens* e) { return (int64_t)e->nx * e->ny * e->nz * e->k * 8; }

// ------------------------------------------------------------------------------ observations
static void obs_free_arrays(mdc_obs* o) {
  cudaFree(o->x); cudaFree(o->y); cudaFree(o->z); cudaFree(o->gid); cudaFree(o->val);
  cudaFree(o->err); cudaFree(o->valid); cudaFree(o->Y); cudaFree(o->ybar); cudaFree(o->Yp);
  cudaFree(o->d);
}

// grow row capacity (and the k-wide arrays once k is known), preserving contents
static int obs_reserve(mdc_obs* o, int64_t cap, int k) {
  mdc_ctx* ctx = o->ctx;
  if (cap <= o->cap && (k == o->k || k == 0)) return MDC_OK;
  const int64_t ncap = (cap > o->cap) ? std::max(cap, o->cap + o->cap / 2) : o->cap;   // geometric growth
  const int nk = k ? k : o->k;
  mdc_obs n = *o;
  auto mv = [&](auto** dst, auto* src, size_t elems_new, size_t elems_old) -> int {
    using T = std::remove_pointer_t<std::remove_pointer_t<decltype(dst)>>;
    if (int rc = dev_alloc<T>(ctx, dst, elems_new)) return rc;
    if (src && elems_old) MDC_CUDA(ctx, cudaMemcpyAsync(*dst, src, elems_old * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return MDC_OK;
  };
  const size_t P = (size_t)o->P;
  if (mv(&n.x, o->x, ncap, P) || mv(&n.y, o->y, ncap, P) || mv(&n.z, o->z, ncap, P) ||
      mv(&n.gid, o->gid, ncap, P) || mv(&n.val, o->val, ncap, P) || mv(&n.err, o->err, ncap, P) ||
      mv(&n.valid, o->valid, ncap, P) || mv(&n.ybar, o->ybar, ncap, o->have_hx ? (size_t)o->P_own : 0) || mv(&n.d, o->d, ncap, o->have_hx ? P : 0))
    return MDC_ERR_CUDA;
  const size_t kold = (o->k == nk && o->have_hx) ? P * (size_t)nk : 0;   // H(x) results exist only after mdc_hx_idw4
  const size_t kown = (o->k == nk && o->have_hx) ? (size_t)o->P_own * (size_t)nk : 0;   // Y, ybar: own rows only (halo rows carry Y')
  if (nk > 0) {
    if (mv(&n.Y, o->Y, (size_t)ncap * nk, kown) || mv(&n.Yp, o->Yp, (size_t)ncap * nk, kold)) return MDC_ERR_CUDA;
  } else {
    n.Y = nullptr; n.Yp = nullptr;
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  obs_free_arrays(o);
  o->x = n.x; o->y = n.y; o->z = n.z; o->gid = n.gid; o->val = n.val; o->err = n.err;
  o->valid = n.valid; o->Y = n.Y; o->ybar = n.ybar; o->Yp = n.Yp; o->d = n.d;
  o->cap = ncap;
  o->k = nk;
  return MDC_OK;
}

int mdc_obs_create(mdc_ctx* ctx, int64_t P, const int32_t* x, const int32_t* y, const int32_t* z,
                   const double* value, const double* err, const uint8_t* valid,
                   const int64_t* gid, mdc_obs** out) {
  if (!ctx || !out) return MDC_ERR_INVALID;
  *out = nullptr;
  if (P < 0 || P > INT32_MAX / 2) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_create: bad P");
  if (P > 0 && (!x || !y || !value || !err)) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_create: null arrays");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  mdc_obs* o = new (std::nothrow) mdc_obs();
  if (!o) MDC_FAIL(ctx, MDC_ERR_INVALID, "out of host memory");
  o->ctx = ctx;
  if (int rc = obs_reserve(o, std::max<int64_t>(P, 1), 0)) { delete o; return rc; }
  o->P = o->P_own = P;
  if (P > 0) {
    std::vector<int32_t> zz;
    if (!z) { zz.assign((size_t)P, 0); z = zz.data(); }
    std::vector<uint8_t> vv;
    if (!valid) { vv.assign((size_t)P, 1); valid = vv.data(); }
    std::vector<int64_t> gg;
    if (!gid) { gg.resize((size_t)P); for (int64_t i = 0; i < P; ++i) gg[(size_t)i] = i; gid = gg.data(); }
    cudaStream_t s = ctx->stream;
    MDC_CUDA(ctx, cudaMemcpyAsync(o->x, x, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->y, y, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->z, z, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->gid, gid, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->val, value, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->err, err, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->valid, valid, P, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaStreamSynchronize(s));  // temporaries above go out of scope
  }
  *out = o;
  return MDC_OK;
}

int mdc_obs_assign(mdc_obs* o, int64_t P, const int32_t* x, const int32_t* y, const int32_t* z,
                   const double* value, const double* err, const uint8_t* valid, const int64_t* gid) {
  mdc_ctx* ctx = o->ctx;
  if (P < 0 || P > INT32_MAX / 2) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_assign: bad P");
  if (P > 0 && (!x || !y || !value || !err)) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_assign: null arrays");
  if (o->geo || o->var) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "mdc_obs_assign: not for geographic / per-variable observation stores");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  o->P = 0;                      // nothing to preserve when growing
  if (int rc = obs_reserve(o, std::max<int64_t>(P, 1), o->k)) return rc;
  o->P = o->P_own = P;
  o->have_hx = false;
  o->index_valid = false;
  if (P > 0) {
    std::vector<int32_t> zz;
    if (!z) { zz.assign((size_t)P, 0); z = zz.data(); }
    std::vector<uint8_t> vv;
    if (!valid) { vv.assign((size_t)P, 1); valid = vv.data(); }
    std::vector<int64_t> gg;
    if (!gid) { gg.resize((size_t)P); for (int64_t i = 0; i < P; ++i) gg[(size_t)i] = i; gid = gg.data(); }
    cudaStream_t s = ctx->stream;
    MDC_CUDA(ctx, cudaMemcpyAsync(o->x, x, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->y, y, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->z, z, P * 4, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->gid, gid, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->val, value, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->err, err, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->valid, valid, P, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaStreamSynchronize(s));  // temporaries above go out of scope
  }
  return MDC_OK;
}

int mdc_obs_destroy(mdc_obs* o) {
  if (!o) return MDC_OK;
  cudaSetDevice(o->ctx->device);
  cudaStreamSynchronize(o->ctx->stream);
  obs_free_arrays(o);
  cudaFree(o->cell_start); cudaFree(o->cell_fill); cudaFree(o->sorted_row);
  cudaFree(o->sx); cudaFree(o->sy); cudaFree(o->sz); cudaFree(o->key);
  cudaFree(o->lat); cudaFree(o->lon); cudaFree(o->lev); cudaFree(o->var); cudaFree(o->qx); cudaFree(o->qy);
  cudaFree(o->cqx); cudaFree(o->cqy); cudaFree(o->slat); cudaFree(o->slon);
  delete o;
  return MDC_OK;
}

int64_t mdc_obs_size(const mdc_obs* o) { return o->P; }
// plain stores: Y'[k], d, value, err, valid, x, y, z, gid; geographic / per-variable stores add lat, lon, level, variable
int mdc_obs_row_doubles(const mdc_obs* o) { return o->k + ((o->geo || o->var) ? 12 : 8); }

// ------------------------------------------------------------------------------ H(x)
int mdc_hx_idw4(mdc_ens* e, mdc_obs* o) {
  mdc_ctx* ctx = e->ctx;
  if (o->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "hx: ens/obs belong to different contexts");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (o->P != o->P_own) {   // a new cycle: drop the halo rows received for the previous one
    o->P = o->P_own;
    o->index_valid = false;
  }
  if (o->geo && !o->located)   // nearest grid point of every observation (IdentityObsOperator.hpp:241-248, 484-530)
    if (int rc = mdc_obs_locate(o, e)) return rc;
  if (o->var && e->nvar == 0) MDC_FAIL(ctx, MDC_ERR_INVALID, "hx: the observations name state variables, the ensemble has none (mdc_ens_set_variables)");
  if (o->var && o->var_max >= e->nvar) MDC_FAIL(ctx, MDC_ERR_INVALID, "hx: an observation observes variable %d, the ensemble has %d", o->var_max, e->nvar);
  if (int rc = obs_reserve(o, o->cap, e->k)) return rc;
  if (o->P == 0) { o->have_hx = true; return MDC_OK; }
  constexpr int W = 8;
  HxGeom g{};
  g.nx = e->nx; g.ny = e->ny; g.nz = e->nz; g.k = e->k; g.gx0 = e->gx0; g.gy0 = e->gy0; g.gnx = e->gnx; g.gny = e->gny;
  g.nzg = e->nvar > 0 ? e->nzg : e->nz; g.nvar = e->nvar;
  for (int v = 0; v < e->nvar; ++v) { g.var_off[v] = e->var_off[v]; g.var_nlev[v] = e->var_nlev[v]; }
  MDC_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
  size_t smem = (size_t)W * e->k * sizeof(double);
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((o->P + W - 1) / W, (int64_t)ctx->sm_count * 8));
  hx_idw4_kernel<W><<<grid, W * 32, smem, ctx->stream>>>(e->X, g, o->P, o->x, o->y, o->z, o->var, o->valid, o->val, o->Y, o->ybar, o->Yp, o->d, ctx->d_flags);
  MDC_LAUNCH_CHECK(ctx);
  int flag2[2] = {0, 0};
  if (int rc = read_small(ctx, ctx->d_flags, flag2, 8)) return rc;
  const int flag = flag2[0];
  if (flag) MDC_FAIL(ctx, MDC_ERR_INVALID, "hx: an observation needs state outside this ensemble's local grid (+halo)");
  o->have_hx = true;
  return MDC_OK;
}

int mdc_hx_download(mdc_obs* o, double* Y, double* ybar, double* Yp, double* d) {
  mdc_ctx* ctx = o->ctx;
  if (!o->have_hx) MDC_FAIL(ctx, MDC_ERR_INVALID, "hx_download: call mdc_hx_idw4 first");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t P = (size_t)o->P, k = (size_t)o->k;
  cudaStream_t s = ctx->stream;
  if (Y) MDC_CUDA(ctx, cudaMemcpyAsync(Y, o->Y, P * k * 8, cudaMemcpyDeviceToHost, s));
  if (Yp) MDC_CUDA(ctx, cudaMemcpyAsync(Yp, o->Yp, P * k * 8, cudaMemcpyDeviceToHost, s));
  if (ybar) MDC_CUDA(ctx, cudaMemcpyAsync(ybar, o->ybar, P * 8, cudaMemcpyDeviceToHost, s));
  if (d) MDC_CUDA(ctx, cudaMemcpyAsync(d, o->d, P * 8, cudaMemcpyDeviceToHost, s));
  MDC_CUDA(ctx, cudaStreamSynchronize(s));
  return MDC_OK;
}

// ------------------------------------------------------------------------------ obs halo rows
__global__ void obs_pack_rows_kernel(int64_t P, int k, const int32_t* x, const int32_t* y,
                                     const int32_t* z, const int64_t* gid, const double* val,
                                     const double* err, const uint8_t* valid, const double* Yp,
                                     const double* d, int ylo, int yhi, double* rows, int64_t cap,
                                     unsigned long long* counter) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rd = k + 8;
  for (int64_t i = warp_global; i < P; i += nwarps) {
    if (y[i] < ylo || y[i] >= yhi) continue;
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((int64_t)slot >= cap) continue;
    double* r = rows + slot * rd;
    for (int j = lane; j < k; j += 32) r[j] = Yp[i * k + j];
    if (lane == 0) {
      r[k + 0] = d[i]; r[k + 1] = val[i]; r[k + 2] = err[i]; r[k + 3] = (double)valid[i];
      r[k + 4] = (double)x[i]; r[k + 5] = (double)y[i]; r[k + 6] = (double)z[i];
      r[k + 7] = (double)gid[i];
    }
  }
}

__global__ void obs_unpack_rows_kernel(int64_t n, int k, const double* rows, int64_t base,
                                       int32_t* x, int32_t* y, int32_t* z, int64_t* gid,
                                       double* val, double* err, uint8_t* valid, double* Yp,
                                       double* d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rd = k + 8;
  for (int64_t i = warp_global; i < n; i += nwarps) {
    const double* r = rows + i * rd;
    const int64_t o = base + i;
    for (int j = lane; j < k; j += 32) Yp[o * k + j] = r[j];
    if (lane == 0) {
      d[o] = r[k + 0]; val[o] = r[k + 1]; err[o] = r[k + 2]; valid[o] = (uint8_t)(r[k + 3] != 0.0);
      x[o] = (int32_t)r[k + 4]; y[o] = (int32_t)r[k + 5]; z[o] = (int32_t)r[k + 6];
      gid[o] = (int64_t)r[k + 7];
    }
  }
}

// Geographic / per-variable stores: rows of k + 12 doubles (.., gid, lat, lon, level, variable).  Own observations
// are selected by grid row (by_box = 0: ylo <= y < yhi) or by a box in the frame of the geography (by_box = 1:
// lat_lo <= lat <= lat_hi and u_lo <= unwrapped (lon - lon_c) <= u_hi) -- a conservative cover of what another rank's
// columns can reach; the haversine test of the selection decides.
struct ObsBox { double lat_lo, lat_hi, u_lo, u_hi, lon_c; };
__global__ void obs_pack_rows_ext_kernel(int64_t P, int k, const int32_t* x, const int32_t* y, const int32_t* z,
                                         const int64_t* gid, const double* val, const double* err, const uint8_t* valid,
                                         const double* Yp, const double* d, const double* lat, const double* lon,
                                         const double* lev, const int32_t* var, int by_box, int ylo, int yhi, ObsBox box,
                                         double* rows, int64_t cap, unsigned long long* counter) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rd = k + 12;
  for (int64_t i = warp_global; i < P; i += nwarps) {
    if (by_box) {
      const double u = (lon[i] - box.lon_c) - 360.0 * rint((lon[i] - box.lon_c) / 360.0);
      if (!(lat[i] >= box.lat_lo && lat[i] <= box.lat_hi && u >= box.u_lo && u <= box.u_hi)) continue;
    } else if (y[i] < ylo || y[i] >= yhi) {
      continue;
    }
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((int64_t)slot >= cap) continue;
    double* r = rows + slot * rd;
    for (int j = lane; j < k; j += 32) r[j] = Yp[i * k + j];
    if (lane == 0) {
      r[k + 0] = d[i]; r[k + 1] = val[i]; r[k + 2] = err[i]; r[k + 3] = (double)valid[i];
      r[k + 4] = (double)x[i]; r[k + 5] = (double)y[i]; r[k + 6] = (double)z[i];
      r[k + 7] = (double)gid[i];
      r[k + 8] = lat ? lat[i] : 0.0; r[k + 9] = lon ? lon[i] : 0.0; r[k + 10] = lev ? lev[i] : 0.0;
      r[k + 11] = var ? (double)var[i] : 0.0;
    }
  }
}

__global__ void obs_unpack_rows_ext_kernel(int64_t n, int k, const double* rows, int64_t base, int32_t* x, int32_t* y,
                                           int32_t* z, int64_t* gid, double* val, double* err, uint8_t* valid, double* Yp,
                                           double* d, double* lat, double* lon, double* lev, int32_t* var) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rd = k + 12;
  for (int64_t i = warp_global; i < n; i += nwarps) {
    const double* r = rows + i * rd;
    const int64_t o = base + i;
    for (int j = lane; j < k; j += 32) Yp[o * k + j] = r[j];
    if (lane == 0) {
      d[o] = r[k + 0]; val[o] = r[k + 1]; err[o] = r[k + 2]; valid[o] = (uint8_t)(r[k + 3] != 0.0);
      x[o] = (int32_t)r[k + 4]; y[o] = (int32_t)r[k + 5]; z[o] = (int32_t)r[k + 6];
      gid[o] = (int64_t)r[k + 7];
      if (lat) { lat[o] = r[k + 8]; lon[o] = r[k + 9]; lev[o] = r[k + 10]; }
      if (var) var[o] = (int32_t)r[k + 11];
    }
  }
}

// grow the arrays only geographic / per-variable stores have, preserving the first P entries
static int obs_reserve_ext(mdc_obs* o, int64_t cap) {
  mdc_ctx* ctx = o->ctx;
  const size_t P = (size_t)o->P;
  auto grow = [&](auto** arr, size_t ncap, bool preserve = true) -> int {
    using T = std::remove_pointer_t<std::remove_pointer_t<decltype(arr)>>;
    T* n = nullptr;
    if (int rc = dev_alloc<T>(ctx, &n, ncap)) return rc;
    if (preserve && *arr && P) MDC_CUDA(ctx, cudaMemcpyAsync(n, *arr, P * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(*arr);
    *arr = n;
    return MDC_OK;
  };
  if (o->geo && (size_t)cap > o->geo_cap) {
    const size_t ncap = std::max((size_t)cap, o->geo_cap + o->geo_cap / 2);
    // (the lattice coordinates qx, qy are rewritten by every index build: nothing to keep)
    if (grow(&o->lat, ncap) || grow(&o->lon, ncap) || grow(&o->lev, ncap) || grow(&o->qx, ncap, false) || grow(&o->qy, ncap, false)) return MDC_ERR_CUDA;
    o->geo_cap = ncap;
  }
  if (o->var && (size_t)cap > o->var_cap) {
    const size_t ncap = std::max((size_t)cap, o->var_cap + o->var_cap / 2);
    if (grow(&o->var, ncap)) return MDC_ERR_CUDA;
    o->var_cap = ncap;
  }
  return MDC_OK;
}

static int obs_pack_impl(mdc_obs* o, int by_box, int ylo, int yhi, ObsBox box, double* dev_rows, int64_t cap, int64_t* n) {
  mdc_ctx* ctx = o->ctx;
  if (!o->have_hx) MDC_FAIL(ctx, MDC_ERR_INVALID, "pack_rows: call mdc_hx_idw4 first");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long* counter = (unsigned long long*)ctx->d_stats;
  MDC_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
  if (o->P_own > 0) {
    obs_pack_rows_ext_kernel<<<grid_for(ctx, o->P_own * 32, 256, 8), 256, 0, ctx->stream>>>(
        o->P_own, o->k, o->x, o->y, o->z, o->gid, o->val, o->err, o->valid, o->Yp, o->d, o->lat, o->lon, o->lev, o->var,
        by_box, ylo, yhi, box, dev_rows, cap, counter);
    MDC_LAUNCH_CHECK(ctx);
  }
  unsigned long long h = 0;
  if (int rc = read_small(ctx, counter, &h, 8)) return rc;
  if (n) *n = (int64_t)h;
  return MDC_OK;
}

int mdc_obs_pack_rows_geo(mdc_obs* o, double lat_lo, double lat_hi, double u_lo, double u_hi, double lon_c,
                          double* dev_rows, int64_t cap, int64_t* n) {
  if (!o->geo) MDC_FAIL(o->ctx, MDC_ERR_INVALID, "pack_rows_geo: the observations carry GRID coordinates (mdc_obs_pack_rows)");
  return obs_pack_impl(o, 1, 0, 0, ObsBox{lat_lo, lat_hi, u_lo, u_hi, lon_c}, dev_rows, cap, n);
}

int mdc_obs_pack_rows(mdc_obs* o, int ylo, int yhi, double* dev_rows, int64_t cap, int64_t* n) {
  mdc_ctx* ctx = o->ctx;
  if (o->geo || o->var) return obs_pack_impl(o, 0, ylo, yhi, ObsBox{0, 0, 0, 0, 0}, dev_rows, cap, n);
  if (!o->have_hx) MDC_FAIL(ctx, MDC_ERR_INVALID, "pack_rows: call mdc_hx_idw4 first");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long* counter = (unsigned long long*)ctx->d_stats;
  MDC_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
  if (o->P_own > 0) {
    obs_pack_rows_kernel<<<grid_for(ctx, o->P_own * 32, 256, 8), 256, 0, ctx->stream>>>(
        o->P_own, o->k, o->x, o->y, o->z, o->gid, o->val, o->err, o->valid, o->Yp, o->d, ylo, yhi, dev_rows, cap, counter);
    MDC_LAUNCH_CHECK(ctx);
  }
  unsigned long long h = 0;
  if (int rc = read_small(ctx, counter, &h, 8)) return rc;
  if (n) *n = (int64_t)h;   // caller compares with cap; rows beyond cap were dropped
  return MDC_OK;
}

int mdc_obs_append_rows(mdc_obs* o, const double* dev_rows, int64_t n) {
  mdc_ctx* ctx = o->ctx;
  if (!o->have_hx) MDC_FAIL(ctx, MDC_ERR_INVALID, "append_rows: call mdc_hx_idw4 first (defines k)");
  if (n <= 0) return MDC_OK;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = obs_reserve(o, o->P + n, o->k)) return rc;
  if (o->geo || o->var) {
    if (o->geo && !o->located) MDC_FAIL(ctx, MDC_ERR_INVALID, "append_rows: locate the store's own observations first");
    if (int rc = obs_reserve_ext(o, o->P + n)) return rc;
    obs_unpack_rows_ext_kernel<<<grid_for(ctx, n * 32, 256, 8), 256, 0, ctx->stream>>>(
        n, o->k, dev_rows, o->P, o->x, o->y, o->z, o->gid, o->val, o->err, o->valid, o->Yp, o->d,
        o->geo ? o->lat : nullptr, o->geo ? o->lon : nullptr, o->geo ? o->lev : nullptr, o->var);
    MDC_LAUNCH_CHECK(ctx);
    o->P += n;
    o->index_valid = false;
    return MDC_OK;
  }
  obs_unpack_rows_kernel<<<grid_for(ctx, n * 32, 256, 8), 256, 0, ctx->stream>>>(
      n, o->k, dev_rows, o->P, o->x, o->y, o->z, o->gid, o->val, o->err, o->valid, o->Yp, o->d);
  MDC_LAUNCH_CHECK(ctx);
  o->P += n;
  o->index_valid = false;
  return MDC_OK;
}

// ------------------------------------------------------------------------------ bucket index
#include "geo_api.inl"

int mdc_obs_index_build(mdc_obs* o, int cell) {
  if (o->geo) MDC_FAIL(o->ctx, MDC_ERR_UNSUPPORTED, "index_build: the index of geographic observations depends on the ensemble's geography and the radius; the query / analysis calls build it");
  return index_build_impl(o, cell);
}

static int index_build_impl(mdc_obs* o, int cell) {
  mdc_ctx* ctx = o->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (cell <= 0) MDC_FAIL(ctx, MDC_ERR_INVALID, "index_build: cell must be > 0");
  const int32_t* kx = o->geo ? o->qx : o->x;   // index coordinates: grid indices, or the lat/lon lattice
  const int32_t* ky = o->geo ? o->qy : o->y;
  cudaStream_t s = ctx->stream;
  const int64_t P = o->P;
  o->cell = cell;
  o->index_P = P;
  int bbox[4] = {0, 0, 0, 0};
  if (P > 0) {
    fill4_kernel<<<1, 32, 0, s>>>(ctx->d_flags + 4, INT_MAX, INT_MAX, INT_MIN, INT_MIN);
    MDC_LAUNCH_CHECK(ctx);
    index_bbox_kernel<<<grid_for(ctx, P, 256, 4), 256, 0, s>>>(kx, ky, P, ctx->d_flags + 4);
    MDC_LAUNCH_CHECK(ctx);
    if (int rc = read_small(ctx, ctx->d_flags + 4, bbox, sizeof(bbox))) return rc;
  }
  // cell boundaries sit at global multiples of `cell` (not at this store's bounding box), so a
  // column sees its candidates in the same (cell, global id) order on every rank / slab
  auto floor_to = [cell](int v) { int q = v / cell; if (v % cell != 0 && v < 0) --q; return q * cell; };
  o->xmin = floor_to(bbox[0]); o->ymin = floor_to(bbox[1]);
  const int64_t ncx = ((int64_t)bbox[2] - o->xmin) / cell + 1, ncy = ((int64_t)bbox[3] - o->ymin) / cell + 1;
  if (ncx * ncy > (1ll << 28)) MDC_FAIL(ctx, MDC_ERR_INVALID, "index_build: %lld x %lld cells is too many; use a larger cell", (long long)ncx, (long long)ncy);
  o->ncx = (int)ncx; o->ncy = (int)ncy;
  const size_t ncell = (size_t)(ncx * ncy);
  if (ncell + 1 > o->cell_cap) {
    cudaFree(o->cell_start); cudaFree(o->cell_fill);
    if (dev_alloc(ctx, &o->cell_start, ncell + 1) || dev_alloc(ctx, &o->cell_fill, ncell)) return MDC_ERR_CUDA;
    o->cell_cap = ncell + 1;
  }
  if ((size_t)P > o->sorted_cap) {
    cudaFree(o->sorted_row); cudaFree(o->sx); cudaFree(o->sy); cudaFree(o->sz); cudaFree(o->key);
    if (dev_alloc(ctx, &o->sorted_row, (size_t)P) || dev_alloc(ctx, &o->sx, (size_t)P) || dev_alloc(ctx, &o->sy, (size_t)P) ||
        dev_alloc(ctx, &o->sz, (size_t)P) || dev_alloc(ctx, &o->key, (size_t)P))
      return MDC_ERR_CUDA;
    o->sorted_cap = (size_t)P;
  }
  MDC_CUDA(ctx, cudaMemsetAsync(o->cell_fill, 0, ncell * sizeof(int32_t), s));
  MDC_CUDA(ctx, cudaMemsetAsync(o->cell_start, 0, (ncell + 1) * sizeof(int32_t), s));
  if (P > 0) {
    // histogram into cell_fill, scan into cell_start, clear cell_fill, scatter, per-cell id sort
    index_key_hist_kernel<<<grid_for(ctx, P, 256, 8), 256, 0, s>>>(kx, ky, P, o->xmin, o->ymin, cell, o->ncx, o->key, o->cell_fill);
    MDC_LAUNCH_CHECK(ctx);
    index_scan_kernel<<<1, 1024, 0, s>>>(o->cell_fill, o->cell_start, (int)ncell);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaMemsetAsync(o->cell_fill, 0, ncell * sizeof(int32_t), s));
    index_scatter_kernel<<<grid_for(ctx, P, 256, 8), 256, 0, s>>>(o->key, P, o->cell_start, o->cell_fill, o->sorted_row);
    MDC_LAUNCH_CHECK(ctx);
    index_cell_sort_kernel<<<mdc_div_up((int64_t)ncell, 128), 128, 0, s>>>(o->cell_start, (int)ncell, o->sorted_row, o->gid, kx, ky, o->z, o->sx, o->sy, o->sz);
    MDC_LAUNCH_CHECK(ctx);
  }
  o->index_valid = true;
  return MDC_OK;
}

static IndexView index_view(const mdc_obs* o, const mdc_ens* e) {
  IndexView iv;
  iv.cell_start = o->cell_start; iv.sorted_row = o->sorted_row;
  iv.sx = o->sx; iv.sy = o->sy; iv.sz = o->sz;
  iv.cell = o->cell; iv.ncx = o->ncx; iv.ncy = o->ncy; iv.xmin = o->xmin; iv.ymin = o->ymin;
  iv.geo = o->geo ? 1 : 0; iv.reach = o->geo_reach;
  iv.cqx = o->cqx; iv.cqy = o->cqy; iv.clat = e->glat; iv.clon = e->glon; iv.slat = o->slat; iv.slon = o->slon;
  iv.levmap = e->levmap;
  return iv;
}

static int ensure_index(mdc_obs* o, mdc_ens* e, double radius) {
  if (o->geo) {
    if (!o->located)   // the index carries the observations' levels
      if (int rc = mdc_obs_locate(o, e)) return rc;
    if (o->index_valid && o->index_P == o->P && o->geo_radius == radius && o->geo_ens == e) return MDC_OK;
    return geo_prepare_index(o, e, radius);
  }
  if (o->index_valid && o->index_P == o->P) return MDC_OK;
  int cell = (int)std::ceil(radius);
  if (cell < 1) cell = 1;
  return mdc_obs_index_build(o, cell);
}

int mdc_obs_index_query_counts(mdc_obs* o, mdc_ens* e, double radius, int32_t* host_counts) {
  mdc_ctx* ctx = o->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = ensure_index(o, e, radius)) return rc;
  const int64_t G = (int64_t)e->nx * e->ny;
  int32_t* dc = nullptr;
  if (dev_alloc(ctx, &dc, (size_t)G)) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemsetAsync(dc, 0xff, G * sizeof(int32_t), ctx->stream));
  const int64_t nown = (int64_t)e->own_nx * e->own_ny;
  auto qk = o->geo ? index_query_counts_kernel<true> : index_query_counts_kernel<false>;
  qk<<<mdc_div_up(nown, 128), 128, 0, ctx->stream>>>(index_view(o, e), e->nx, e->own_nx, e->own_ny, e->gx0, e->gy0, radius, dc);
  MDC_LAUNCH_CHECK(ctx);
  MDC_CUDA(ctx, cudaMemcpyAsync(host_counts, dc, G * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(dc);
  return MDC_OK;
}

int mdc_obs_index_query_lists(mdc_obs* o, mdc_ens* e, double radius, const int64_t* cols,
                              int64_t ncols, int32_t cap, int64_t* host_lists,
                              int32_t* host_counts) {
  mdc_ctx* ctx = o->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ncols <= 0 || cap <= 0) return MDC_OK;
  if (int rc = ensure_index(o, e, radius)) return rc;
  int64_t *dcols = nullptr, *dl = nullptr;
  int32_t* dc = nullptr;
  if (dev_alloc(ctx, &dcols, (size_t)ncols) || dev_alloc(ctx, &dl, (size_t)ncols * cap) || dev_alloc(ctx, &dc, (size_t)ncols)) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemcpyAsync(dcols, cols, ncols * 8, cudaMemcpyHostToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaMemsetAsync(dl, 0xff, (size_t)ncols * cap * 8, ctx->stream));
  auto qk = o->geo ? index_query_lists_kernel<true> : index_query_lists_kernel<false>;
  qk<<<mdc_div_up(ncols, 64), 64, 0, ctx->stream>>>(index_view(o, e), o->gid, e->nx, e->gx0, e->gy0, radius, dcols, ncols, cap, dl, dc);
  MDC_LAUNCH_CHECK(ctx);
  MDC_CUDA(ctx, cudaMemcpyAsync(host_lists, dl, (size_t)ncols * cap * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaMemcpyAsync(host_counts, dc, ncols * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(dcols); cudaFree(dl); cudaFree(dc);
  return MDC_OK;
}

// ------------------------------------------------------------------------------ LETKF
static int letkf_launch(mdc_ens* e, mdc_obs* o, const mdc_letkf_params* p, const long long* dcols,
                        long long ncols, double* dW, long long w_col) {
  mdc_ctx* ctx = e->ctx;
  const int k = e->k;
  if (k > 128) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: k=%d > 128 members not supported", k);
  if (k < 2) MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: needs at least 2 members");
  if (p->mode < 0 || p->mode > 2) MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: bad mode");
  if (p->loc < 0 || p->loc > MDC_LOC_REF_GASPARI_COHN) MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: bad localisation function");
  if (!(p->inflation > 0.0)) MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: inflation must be > 0");
  // geographic observations / multi-variable states run on the EXT instantiations of the column kernels
  const bool ext = o->geo || e->levmap != nullptr;
  if (ext && p->solver == MDC_SOLVER_NEWTON_SCHULZ_FULL)
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: geographic observations and multi-variable states run on the AUTO, JACOBI or NEWTON_SCHULZ solvers");
  const size_t smem = lk_smem_bytes(k, p->mode);
  if ((int)smem > ctx->max_smem_optin)
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: k=%d mode=%d needs %zu B shared memory > %d available", k, p->mode, smem, ctx->max_smem_optin);
  ColParams cp;
  cp.X = e->X;
  if (!e->mean) MDC_CUDA(ctx, cudaMalloc((void**)&e->mean, (size_t)e->nx * e->ny_cap * e->nz * sizeof(double)));
  cp.mean_out = e->mean;
  cp.nx = e->nx; cp.ny = e->ny; cp.nz = e->nz; cp.k = k; cp.own_nx = e->own_nx; cp.own_ny = e->own_ny;
  cp.gx0 = e->gx0; cp.gy0 = e->gy0;
  cp.iv = index_view(o, e);
  cp.Yp = o->Yp; cp.d = o->d; cp.err = o->err; cp.valid = o->valid;
  cp.radius = p->radius; cp.radius_v = p->radius_v; cp.inflation = p->inflation;
  cp.mode = p->mode; cp.loc = p->loc; cp.use_R = p->use_R;
  cp.loc_scale = p->loc_scale > 0.0 ? p->loc_scale : p->radius;
  cp.loc_scale_v = p->radius_v > 0.0 && p->radius > 0.0 ? p->radius_v * (cp.loc_scale / p->radius) : 1.0;
  cp.kappa_max = p->kappa_max > 0.0 ? std::min(p->kappa_max, (double)NSP_KAPPA_TABLE_MAX) : (double)NSP_KAPPA_MAX_DEFAULT;
  cp.max_sweeps = p->max_sweeps > 0 ? p->max_sweeps : 40;
  cp.jtol = p->jacobi_tol > 0.0 ? p->jacobi_tol : 1e-11;
  cp.stats = ctx->d_stats;
  cp.W_out = dW; cp.w_col = w_col;
  cp.cols = dcols; cp.ncols = ncols;
  const long long total_cols = dcols ? ncols : (long long)e->own_nx * e->own_ny;
  const int sms = std::max(1, ctx->sm_count - std::max(0, std::min(p->sm_reserve, ctx->sm_count / 2)));
  cp.redo_items = nullptr; cp.redo_count = nullptr; cp.redo_consume = 0;
  cp.small_items = nullptr; cp.small_count = nullptr;
  cp.work_items = nullptr; cp.work_count = nullptr; cp.work_consume = 0;
  if (p->mode == MDC_MODE_CANONICAL && p->solver == MDC_SOLVER_NEWTON_SCHULZ && (k < 24 || k > 128))
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: the Newton-Schulz solver supports 24 <= k <= 128 (k=%d)", k);
  if (p->mode == MDC_MODE_CANONICAL && p->solver == MDC_SOLVER_NEWTON_SCHULZ_FULL && (k < 24 || k > 80))
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: the full-product Newton-Schulz solver supports 24 <= k <= 80 (k=%d)", k);
  // full-product Newton-Schulz (letkf_ns.cuh): four padded k x k buffers, one CTA per SM
  auto launch_ns_full = [&](const ColParams& cq, long long work) -> int {
    const int lch = std::min(32, (e->nz + 7) & ~7);   // levels per update chunk (multiple of 8)
    const size_t smem3 = ns_smem_bytes(k, lch);
    if ((int)smem3 > ctx->max_smem_optin)
      MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: k=%d needs %zu B shared memory > %d available", k, smem3, ctx->max_smem_optin);
    auto launch3 = [&](auto kern) -> int {
      MDC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      int grid = (int)std::max<long long>(1, std::min<long long>(work, (long long)sms));
      kern<<<grid, NS_THREADS, smem3, ctx->stream>>>(cq, lch);
      MDC_LAUNCH_CHECK(ctx);
      return MDC_OK;
    };
    if (k <= 32) return launch3(letkf_ns_kernel<2>);
    if (k <= 48) return launch3(letkf_ns_kernel<3>);
    if (k <= 64) return launch3(letkf_ns_kernel<4>);
    return launch3(letkf_ns_kernel<5>);
  };
  // blocked Jacobi eigensolver (letkf_v2.cuh): level-chunk sized so that two CTAs fit one SM when k allows
  auto launch_jacobi = [&](const ColParams& cq, long long work) -> int {
    const int lchmax = std::max(4, 2560 / k);
    const int nchunk = (e->nz + lchmax - 1) / lchmax;
    const int lch = (e->nz + nchunk - 1) / nchunk;
    const size_t smem2 = v2_smem_bytes(k, lch);
    if ((int)smem2 > ctx->max_smem_optin)
      MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: k=%d needs %zu B shared memory > %d available", k, smem2, ctx->max_smem_optin);
    auto launch2 = [&](auto kern, int nt) -> int {
      MDC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      int occ = 1;
      MDC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem2));
      if (occ < 1) occ = 1;
      int grid = (int)std::max<long long>(1, std::min<long long>(work, (long long)sms * occ));
      kern<<<grid, nt, smem2, ctx->stream>>>(cq, lch);
      MDC_LAUNCH_CHECK(ctx);
      return MDC_OK;
    };
    // <NT, MINB, LG, RPL, TM, TMY>: NT = (k/8 block pairs) * LG lanes rounded up to whole warps
#define MDC_JAC(NT, ...) (ext ? launch2(letkf_canonical_kernel<NT, __VA_ARGS__, true>, NT) : launch2(letkf_canonical_kernel<NT, __VA_ARGS__>, NT))
    if (k <= 16) return MDC_JAC(32, 8, 4, 4, 1, 8);
    if (k <= 24) return MDC_JAC(32, 8, 4, 6, 2, 12);
    if (k <= 40) return MDC_JAC(64, 4, 8, 5, 3, 10);
    if (k <= 64) return MDC_JAC(64, 4, 8, 8, 4, 16);
    if (k <= 80) return MDC_JAC(160, 2, 16, 5, 5, 8);
    return MDC_JAC(256, 1, 16, 8, 8, 8);
#undef MDC_JAC
  };
  const bool v1 = getenv("MDC_LETKF_V1") != nullptr;
  if (p->mode == MDC_MODE_CANONICAL && p->solver == MDC_SOLVER_NEWTON_SCHULZ_FULL && !v1)
    return launch_ns_full(cp, total_cols);
  if (p->mode == MDC_MODE_CANONICAL && p->solver != MDC_SOLVER_JACOBI && k >= 24 && k <= 128 && !v1) {
    // packed symmetric Newton-Schulz (letkf_nsp.cuh), then its redo list through the full-product
    // kernel (k <= 80) or the Jacobi kernel
    const int nxf = p->radius_v > 0.0 ? e->nz : 1;
    const size_t need = (size_t)total_cols * nxf;
    if (ctx->redo_cap < need) {      // three lists: redo (ill-conditioned), small (few local observations), work
      if (ctx->redo_items) cudaFree(ctx->redo_items);
      ctx->redo_items = nullptr; ctx->redo_cap = 0;
      MDC_CUDA(ctx, cudaMalloc((void**)&ctx->redo_items, 3 * need * sizeof(long long)));
      ctx->redo_cap = need;
    }
    cp.redo_items = ctx->redo_items;
    cp.redo_count = reinterpret_cast<unsigned*>(ctx->d_flags + 12);
    MDC_CUDA(ctx, cudaMemsetAsync(ctx->d_flags + 12, 0, 3 * sizeof(unsigned), ctx->stream));
    const bool smallp = !getenv("MDC_LETKF_NO_SMALLP");
    auto launch_smallp = [&](auto kern, const ColParams& cq, int pmax = SP_PMAX, int warps = SP_WARPS) -> int {
      const size_t smems = smallp_smem_bytes(pmax, warps);
      MDC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems));
      int occ = 1;
      MDC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smems));
      kern<<<sms * std::max(1, occ), warps * 32, smems, ctx->stream>>>(cq);
      MDC_LAUNCH_CHECK(ctx);
      return MDC_OK;
    };
    // Per-level analyses: a first pass (one warp per column) finishes the transforms with no or few local
    // observations in observation space and lists the others for the k-space kernel.  Otherwise the k-space
    // kernel defers its small transforms to the observation-space kernel after its own selection.
    const bool classify = smallp && nxf > 1 && !dW;
    if (classify) {
      cp.work_items = ctx->redo_items + 2 * ctx->redo_cap;
      cp.work_count = reinterpret_cast<unsigned*>(ctx->d_flags + 14);
      // first pass (p <= 24), then the transforms with 24 < p <= 32 from its second list: both in observation space
      ColParams cc = cp;
      cc.small_items = ctx->redo_items + ctx->redo_cap;
      cc.small_count = reinterpret_cast<unsigned*>(ctx->d_flags + 13);
      if (ext) {          // (geographic observations / multi-variable states: haversine selection, levels inside variables)
        if (int rc = launch_smallp(letkf_smallp_classify_kernel<SP_WARPS_P1, true>, cc, SP_PMAX, SP_WARPS_P1)) return rc;
        if (int rc = launch_smallp(letkf_smallp_kernel<true, SP_PMAX2, SP_WARPS_P2>, cc, SP_PMAX2, SP_WARPS_P2)) return rc;
      } else {
        if (int rc = launch_smallp(letkf_smallp_classify_kernel<SP_WARPS_P1>, cc, SP_PMAX, SP_WARPS_P1)) return rc;
        if (int rc = launch_smallp(letkf_smallp_kernel<false, SP_PMAX2, SP_WARPS_P2>, cc, SP_PMAX2, SP_WARPS_P2)) return rc;
      }
      cp.work_consume = 1;
    } else if (smallp) {
      cp.small_items = ctx->redo_items + ctx->redo_cap;
      cp.small_count = reinterpret_cast<unsigned*>(ctx->d_flags + 13);
    }
    const int lch = nsp_level_chunk(k, e->nz);
    // the kernel's instantiations live in their own translation units (nsp_tu.cu, nsp_launch.h)
    const int ntile = std::min((k + 7) >> 3, 16);   // tile rows: the kernel is specialised on the exact count
    int rc = NSP_NOT_MINE;
#define MDC_NSP_TRY(LO, HI) \
    if (rc == NSP_NOT_MINE) rc = nsp_launch_##LO##_##HI(ntile, &cp, lch, sms, ext ? 1 : 0, cp.work_consume, total_cols, ctx);
    NSP_GROUPS(MDC_NSP_TRY)
#undef MDC_NSP_TRY
    if (rc == NSP_NOT_MINE) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "letkf: no packed kernel for k=%d", k);
    if (rc) return rc;
    if (cp.small_items)
      if (int rc2 = ext ? launch_smallp(letkf_smallp_kernel<true>, cp) : launch_smallp(letkf_smallp_kernel<false>, cp)) return rc2;
    ColParams cq = cp;
    cq.redo_consume = 1;
    cq.work_consume = 0;
    return (k <= 80 && !ext) ? launch_ns_full(cq, (long long)need) : launch_jacobi(cq, (long long)need);
  }
  if (p->mode == MDC_MODE_CANONICAL && !v1) return launch_jacobi(cp, total_cols);
  const int nr = (k + 31) / 32;
  auto launch = [&](auto kern) -> int {
    MDC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    MDC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LK_THREADS, smem));
    if (occ < 1) occ = 1;
    int grid = (int)std::max<long long>(1, std::min<long long>(total_cols, (long long)sms * occ));
    kern<<<grid, LK_THREADS, smem, ctx->stream>>>(cp);
    MDC_LAUNCH_CHECK(ctx);
    return MDC_OK;
  };
  if (ext) switch (nr) {
    case 1: return launch(letkf_column_kernel<1, true>);
    case 2: return launch(letkf_column_kernel<2, true>);
    case 3: return launch(letkf_column_kernel<3, true>);
    default: return launch(letkf_column_kernel<4, true>);
  }
  switch (nr) {
    case 1: return launch(letkf_column_kernel<1>);
    case 2: return launch(letkf_column_kernel<2>);
    case 3: return launch(letkf_column_kernel<3>);
    default: return launch(letkf_column_kernel<4>);
  }
}

int mdc_letkf_analyse(mdc_ens* e, mdc_obs* o, const mdc_letkf_params* p, mdc_letkf_stats* st) {
  mdc_ctx* ctx = e->ctx;
  if (o->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: ens/obs belong to different contexts");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  MDC_CUDA(ctx, cudaEventRecord(ctx->pe[0], s));
  // K2/K3: Y' = H(X) - mean once from the background ensemble (snapshot semantics), unless the
  // caller already ran H (multi-GPU: H, then halo exchange, then analyse)
  if (!o->have_hx) {
    if (int rc = mdc_hx_idw4(e, o)) return rc;
  } else if (o->k != e->k) {
    MDC_FAIL(ctx, MDC_ERR_INVALID, "letkf: Y' has k=%d, ensemble has k=%d", o->k, e->k);
  }
  MDC_CUDA(ctx, cudaEventRecord(ctx->pe[1], s));
  if (int rc = ensure_index(o, e, p->radius)) return rc;
  MDC_CUDA(ctx, cudaEventRecord(ctx->pe[2], s));
  MDC_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, 16 * sizeof(long long), s));
  if (int rc = letkf_launch(e, o, p, nullptr, 0, nullptr, -1)) return rc;
  MDC_CUDA(ctx, cudaEventRecord(ctx->pe[3], s));
  long long h[16];
  if (int rc = read_small(ctx, ctx->d_stats, h, sizeof(h))) return rc;
  o->have_hx = false;  // the ensemble changed: Y' is stale for a next cycle
  if (st) {
    memset(st, 0, sizeof(*st));
    cudaEventElapsedTime(&st->ms_hx, ctx->pe[0], ctx->pe[1]);
    cudaEventElapsedTime(&st->ms_index, ctx->pe[1], ctx->pe[2]);
    cudaEventElapsedTime(&st->ms_columns, ctx->pe[2], ctx->pe[3]);
    cudaEventElapsedTime(&st->ms_total, ctx->pe[0], ctx->pe[3]);
    st->columns = h[5];
    st->sum_local_obs = h[0];
    st->max_local_obs = (int32_t)h[1];
    st->sum_sweeps = h[2];
    st->max_sweeps = (int32_t)h[3];
    st->numeric_failures = (int32_t)h[4];
    st->redo_transforms = (int32_t)h[6];
    st->small_transforms = h[7];
  }
  if (h[4]) MDC_FAIL(ctx, MDC_ERR_NUMERIC, "letkf: %lld column transforms failed (non-SPD matrix); those columns were left unchanged", h[4]);
  return MDC_OK;
}

int mdc_letkf_column_transform(mdc_ens* e, mdc_obs* o, const mdc_letkf_params* p, int64_t col,
                               double* host_W) {
  // Debug/test entry: runs the column kernel on ONE column and returns its W. The column's state
  // IS updated (callers use a scratch ensemble).
  mdc_ctx* ctx = e->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!o->have_hx) { if (int rc = mdc_hx_idw4(e, o)) return rc; }
  if (int rc = ensure_index(o, e, p->radius)) return rc;
  double* dW = nullptr;
  long long* dcol = nullptr;
  if (dev_alloc(ctx, &dW, (size_t)e->k * e->k) || dev_alloc(ctx, &dcol, 1)) return MDC_ERR_CUDA;
  long long c = col;
  MDC_CUDA(ctx, cudaMemcpyAsync(dcol, &c, 8, cudaMemcpyHostToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, 16 * sizeof(long long), ctx->stream));
  int rc = letkf_launch(e, o, p, dcol, 1, dW, c);
  if (!rc) {
    MDC_CUDA(ctx, cudaMemcpyAsync(host_W, dW, (size_t)e->k * e->k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(dW); cudaFree(dcol);
  return rc;
}

// ------------------------------------------------------------------------------ verification metrics
int mdc_ens_metrics(mdc_ens* e, mdc_ens* truth, mdc_metrics* out, double* host_spread) {
  mdc_ctx* ctx = e->ctx;
  if (truth->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "metrics: ensemble and truth belong to different contexts");
  if (truth->k != 1 || truth->nx != e->nx || truth->ny != e->ny || truth->nz != e->nz)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "metrics: truth must be a one-member ensemble on the same %dx%dx%d grid", e->nx, e->ny, e->nz);
  if (e->k < 2 || e->k > 128) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "metrics: needs 2 <= k <= 128 members (k=%d)", e->k);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t G = (int64_t)e->nx * e->ny, npoints = G * e->nz;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((npoints + 7) / 8, (int64_t)ctx->sm_count * 8));
  double *partial = nullptr, *res = nullptr, *spread = nullptr, *spread_h = nullptr;
  if (dev_alloc(ctx, &partial, (size_t)grid * MT_NSUM) || dev_alloc(ctx, &res, 8)) return MDC_ERR_CUDA;
  if (host_spread && (dev_alloc(ctx, &spread, (size_t)npoints) || dev_alloc(ctx, &spread_h, (size_t)npoints))) return MDC_ERR_CUDA;
  metrics_points_kernel<<<grid, 256, 0, ctx->stream>>>(e->X, truth->X, npoints, e->k, spread, partial);
  MDC_LAUNCH_CHECK(ctx);
  metrics_final_kernel<<<1, 32, 0, ctx->stream>>>(partial, grid, (double)npoints, res);
  MDC_LAUNCH_CHECK(ctx);
  double h[5];
  MDC_CUDA(ctx, cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  if (host_spread) {
    mean_to_host_order_kernel<<<mdc_div_up(npoints, 256), 256, 0, ctx->stream>>>(spread, spread_h, G, e->nz);
    MDC_LAUNCH_CHECK(ctx);
    MDC_CUDA(ctx, cudaMemcpyAsync(host_spread, spread_h, (size_t)npoints * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->rmse = h[0]; out->bias = h[1]; out->correlation = h[2]; out->crps = h[3]; out->avg_spread = h[4];
  cudaFree(partial); cudaFree(res);
  if (spread) cudaFree(spread);
  if (spread_h) cudaFree(spread_h);
  return MDC_OK;
}

// ------------------------------------------------------------------------------ microbenchmarks
int mdc_bench_fp64_fma(mdc_ctx* ctx, double* tflops) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  double* d = nullptr;
  if (dev_alloc(ctx, &d, 8)) return MDC_ERR_CUDA;
  const int grid = ctx->sm_count * 8;
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    mdc_timer_start(ctx);
    mb_fp64_fma_kernel<<<grid, 256, 0, ctx->stream>>>(d, 1.0000001, 1e-9);
    MDC_LAUNCH_CHECK(ctx);
    float ms;
    mdc_timer_stop(ctx, &ms);
    if (it > 0) best = std::min(best, ms);
  }
  *tflops = 2.0 * (double)grid * 256 * MB_ITERS * MB_ILP / (best * 1e-3) / 1e12;
  cudaFree(d);
  return MDC_OK;
}

int mdc_bench_fp64_dmma(mdc_ctx* ctx, double* tflops) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  double* d = nullptr;
  if (dev_alloc(ctx, &d, 8)) return MDC_ERR_CUDA;
  const int grid = ctx->sm_count * 8;
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    mdc_timer_start(ctx);
    mb_fp64_dmma_kernel<<<grid, 256, 0, ctx->stream>>>(d, 1.0000001, 1e-9);
    MDC_LAUNCH_CHECK(ctx);
    float ms;
    mdc_timer_stop(ctx, &ms);
    if (it > 0) best = std::min(best, ms);
  }
  // one m8n8k4 mma per warp = 8*8*4 FMAs = 512 flops
  *tflops = 512.0 * (double)grid * (256 / 32) * MB_ITERS * MB_ILP / (best * 1e-3) / 1e12;
  cudaFree(d);
  return MDC_OK;
}

int mdc_bench_hbm_copy(mdc_ctx* ctx, double* gbs) {
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = (1ll << 30) / sizeof(double2);  // 1 GiB each way
  double2 *a = nullptr, *b = nullptr;
  if (dev_alloc(ctx, &a, (size_t)n) || dev_alloc(ctx, &b, (size_t)n)) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemsetAsync(a, 0, n * sizeof(double2), ctx->stream));
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    mdc_timer_start(ctx);
    mb_copy_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(a, b, n);
    MDC_LAUNCH_CHECK(ctx);
    float ms;
    mdc_timer_stop(ctx, &ms);
    if (it > 0) best = std::min(best, ms);
  }
  *gbs = 2.0 * (double)n * sizeof(double2) / (best * 1e-3) / 1e9;
  cudaFree(a); cudaFree(b);
  return MDC_OK;
}

}  // extern "C"

#include "global_api.inl"
#include "lwenkf_api.inl"
