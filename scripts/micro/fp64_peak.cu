// Micro-benchmark: achievable FP64 FMA and DMMA (mma.sync.m8n8k4.f64) rates per SM on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fma_kernel(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double acc[16][2];
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3;
  double a = 1.0 + threadIdx.x * 1e-6, b = 0.5;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32, iters = 20000;
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) fma_kernel<<<sms, threads>>>(out, iters);
        else dmma_kernel<<<sms, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      const double fmas = which == 0 ? (double)sms * threads * iters * 16 : (double)sms * warps * iters * 16 * 256;
      printf("%s warps/SM=%2d  %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", which ? "DMMA" : "DFMA",
             warps, best, 2 * fmas / best / 1e9, fmas / (best * 1e-3) / sms / (khz * 1e3), khz / 1000);
    }
  }
  return 0;
}
