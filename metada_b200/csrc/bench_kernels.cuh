// Microbenchmarks that measure this B200's FP64 denominators for the roofline (the driver's
// MEASURED_PEAKS.json only carries HBM copy and bf16 GEMM): DFMA pipe, DMMA (mma.sync m8n8k4 f64).
#pragma once
#include "mdc_internal.cuh"

#define MB_ITERS 4096
#define MB_ILP 8

__global__ void __launch_bounds__(256) mb_fp64_fma_kernel(double* out, double a, double b) {
  double acc[MB_ILP];
#pragma unroll
  for (int i = 0; i < MB_ILP; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < MB_ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < MB_ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) mb_fp64_dmma_kernel(double* out, double a, double b) {
  double c0[MB_ILP], c1[MB_ILP];
#pragma unroll
  for (int i = 0; i < MB_ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < MB_ILP; ++i) {
      asm volatile(
          "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
          : "+d"(c0[i]), "+d"(c1[i])
          : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < MB_ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

__global__ void mb_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = src[e];
}
