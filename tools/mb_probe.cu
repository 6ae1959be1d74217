// mb_probe.cu -- standalone pipe microbenchmarks that shape the column kernel's design (round 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/mb_probe tools/mb_probe.cu
// Prints one JSON object.  Questions answered:
//   1. are the FP64 FMA pipe and the FP64 tensor (DMMA) sub-pipe independent, i.e. does a mix of DFMA and
//      DMMA exceed either alone?  (in one warp's instruction stream, and on separate warps)
//   2. how fast is DMMA when its fragments come from shared memory (2, 1, 0.5 LDS.64 per DMMA)?
//   3. legacy mma.sync rates for tf32 / bf16 (what a low-precision seed could use)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define ITERS 2048
#define ILP 8

#define DMMA(c0, c1, a, b)                                                                        \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" \
               : "+d"(c0), "+d"(c1)                                                               \
               : "d"(a), "d"(b))

// mode 0: DFMA only, 1: DMMA only, 2: both in every warp (ILP DMMA + nf * ILP DFMA per iteration),
// 3: even warps DMMA, odd warps DFMA
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b) {
  double c0[ILP], c1[ILP], f[ILP * (NF > 0 ? NF : 1)];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < ILP * (NF > 0 ? NF : 1); ++i) f[i] = (double)(threadIdx.x + i);
  const bool dm = (MODE == 1) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
  const bool fm = (MODE == 0) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
  for (int it = 0; it < ITERS; ++it) {
    if (dm) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) DMMA(c0[i], c1[i], a, b);
    }
    if (fm) {
#pragma unroll
      for (int i = 0; i < ILP * (NF > 0 ? NF : 1); ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < ILP * (NF > 0 ? NF : 1); ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

// DMMA fed from shared memory.  LPD2 = LDS.64 per DMMA times 2 (4: a and b fresh for every MMA; 2: b fresh, a
// reused across the ILP tiles; 1: one fragment per two MMAs)
template <int LPD2>
__global__ void __launch_bounds__(256) k_dmma_smem(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  const double* base = sm + warp * 64 + lane;
  for (int it = 0; it < iters; ++it) {
    const double* p = base + ((it & 7) << 9);
    if (LPD2 == 4) {
      double a[ILP], b[ILP];
#pragma unroll
      for (int i = 0; i < ILP; ++i) { a[i] = p[i * 32]; b[i] = p[4096 + i * 32]; }
#pragma unroll
      for (int i = 0; i < ILP; ++i) DMMA(c0[i], c1[i], a[i], b[i]);
    } else if (LPD2 == 2) {
      double a = p[0], b[ILP];
#pragma unroll
      for (int i = 0; i < ILP; ++i) b[i] = p[4096 + i * 32];
#pragma unroll
      for (int i = 0; i < ILP; ++i) DMMA(c0[i], c1[i], a, b[i]);
    } else {
      double a[2], b[ILP / 2];
      a[0] = p[0]; a[1] = p[32];
#pragma unroll
      for (int i = 0; i < ILP / 2; ++i) b[i] = p[4096 + i * 32];
#pragma unroll
      for (int i = 0; i < ILP; ++i) DMMA(c0[i], c1[i], a[i & 1], b[i >> 1]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// legacy tensor path: tf32 m16n8k8 and bf16 m16n8k16, register operands
template <int KIND>
__global__ void __launch_bounds__(256) k_hmma(float* out, unsigned a, unsigned b) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a), "r"(a), "r"(a), "r"(a), "r"(b), "r"(b));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a), "r"(a), "r"(a), "r"(a), "r"(b), "r"(b));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

template <typename F>
static float time_ms(F&& launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  double* d; cudaMalloc(&d, 64);
  const double a = 1.0000001, b = 1e-9;
  const int grid = sms * 8;
  const double thr = (double)grid * 256, wrp = (double)grid * 8;
  printf("{\"sms\": %d", sms);
  float ms;
  ms = time_ms([&] { k_mix<0, 1><<<grid, 256>>>(d, a, b); });
  printf(", \"dfma_only_tflops\": %.2f", 2.0 * thr * ITERS * ILP / ms / 1e9);
  ms = time_ms([&] { k_mix<1, 0><<<grid, 256>>>(d, a, b); });
  printf(", \"dmma_only_tflops\": %.2f", 512.0 * wrp * ITERS * ILP / ms / 1e9);
  // in-warp mix: ILP DMMA (512 flop per warp each) + NF*ILP DFMA (64 flop per warp each)
  ms = time_ms([&] { k_mix<2, 1><<<grid, 256>>>(d, a, b); });
  printf(", \"inwarp_mix_nf1\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", ms,
         512.0 * wrp * ITERS * ILP / ms / 1e9, 2.0 * thr * ITERS * ILP / ms / 1e9);
  ms = time_ms([&] { k_mix<2, 4><<<grid, 256>>>(d, a, b); });
  printf(", \"inwarp_mix_nf4\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", ms,
         512.0 * wrp * ITERS * ILP / ms / 1e9, 2.0 * thr * ITERS * ILP * 4 / ms / 1e9);
  ms = time_ms([&] { k_mix<2, 8><<<grid, 256>>>(d, a, b); });
  printf(", \"inwarp_mix_nf8\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", ms,
         512.0 * wrp * ITERS * ILP / ms / 1e9, 2.0 * thr * ITERS * ILP * 8 / ms / 1e9);
  // warp split: half the warps each; per-half work chosen so both finish at about the same time if the pipes are independent
  ms = time_ms([&] { k_mix<3, 8><<<grid, 256>>>(d, a, b); });
  printf(", \"warp_split_nf8\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", ms,
         512.0 * (wrp / 2) * ITERS * ILP / ms / 1e9, 2.0 * (thr / 2) * ITERS * ILP * 8 / ms / 1e9);
  // DMMA from shared memory, 1 and 2 CTAs (of 8 warps) per SM
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int g2 = sms * ctas, it2 = 16384;
    const double w2 = (double)g2 * 8;
    cudaFuncSetAttribute(k_dmma_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(k_dmma_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(k_dmma_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    ms = time_ms([&] { k_dmma_smem<4><<<g2, 256, 96 * 1024>>>(d, it2); });
    printf(", \"dmma_smem_2lds_per_mma_%dcta\": %.2f", ctas, 512.0 * w2 * it2 * ILP / ms / 1e9);
    ms = time_ms([&] { k_dmma_smem<2><<<g2, 256, 96 * 1024>>>(d, it2); });
    printf(", \"dmma_smem_1.125lds_per_mma_%dcta\": %.2f", ctas, 512.0 * w2 * it2 * ILP / ms / 1e9);
    ms = time_ms([&] { k_dmma_smem<1><<<g2, 256, 96 * 1024>>>(d, it2); });
    printf(", \"dmma_smem_0.75lds_per_mma_%dcta\": %.2f", ctas, 512.0 * w2 * it2 * ILP / ms / 1e9);
  }
  float* df; cudaMalloc(&df, 64);
  ms = time_ms([&] { k_hmma<0><<<grid, 256>>>(df, 0x3f800000u, 0x3f800000u); });
  printf(", \"mma_sync_tf32_m16n8k8_tflops\": %.1f", 2.0 * 16 * 8 * 8 * wrp * ITERS * ILP / ms / 1e9);
  ms = time_ms([&] { k_hmma<1><<<grid, 256>>>(df, 0x3f803f80u, 0x3f803f80u); });
  printf(", \"mma_sync_bf16_m16n8k16_tflops\": %.1f", 2.0 * 16 * 8 * 16 * wrp * ITERS * ILP / ms / 1e9);
  cudaError_t e = cudaDeviceSynchronize();
  printf(", \"cuda\": \"%s\"}\n", cudaGetErrorString(e));
  return 0;
}
