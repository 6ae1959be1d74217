// mb_issue.cu -- how many warps per SM sub-partition does FP64 DMMA need to saturate its pipe, with and without the
// shared-memory loads of the column kernel's product loop in the instruction stream?  (round 2: the per-warp
// product times of letkf_nsp_kernel differ by 30 % between the two warps of a sub-partition.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mb_issue tools/mb_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c0, c1, a, b)                                                                        \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" \
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
__device__ __forceinline__ double lds_v(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// MODE 0: register operands; 1: ring of three B fragments requested two MMAs ahead + one A fragment per 7 MMAs (the
// product loop's pattern); ILP accumulators
template <int MODE, int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 8 + (threadIdx.x >> 5) * 512;
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  double a = 1.0000001, br[3] = {1e-9, 2e-9, 3e-9};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
      const unsigned p = base + ((it & 7) << 12);
      a = lds_v(p);
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        br[(i + 2) % 3] = lds_v(p + 256 * (i + 1));
        DMMA(c0[i], c1[i], a, br[i % 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < ILP; ++i) DMMA(c0[i], c1[i], a, br[i % 3]);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE, int ILP>
static void run(const char* name, int sms, int warps, double* d, long long* dc) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<MODE, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // one CTA per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<MODE, ILP><<<sms, warps * 32, 200 * 1024>>>(d, iters, dc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf(" {\"case\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"tflops\": %.2f, \"cycles_per_dmma_per_warp\": %.1f},\n", name, warps, ILP,
         512.0 * sms * warps * iters * ILP / best / 1e9, (double)c / ((double)iters * ILP));
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  double* d; cudaMalloc(&d, 64);
  long long* dc; cudaMalloc(&dc, 64);
  printf("[\n");
  for (int w : {4, 8, 12, 16, 32}) {
    run<0, 7>("reg", sms, w, d, dc);
    run<1, 7>("lds_ring", sms, w, d, dc);
    run<1, 4>("lds_ring", sms, w, d, dc);
  }
  printf(" {\"cuda\": \"%s\"}\n]\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
