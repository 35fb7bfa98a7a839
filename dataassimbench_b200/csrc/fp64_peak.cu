// Measured FP64 roofline denominators (same kernels as tools/fp64_peak.cu, callable through the
// C ABI so bench.py can state the DFMA / DMMA peaks of the very device it timed).
#include "common.cuh"
#include "kernels.h"

namespace dab {

__global__ void __launch_bounds__(1024) peak_dfma_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += acc[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024) peak_dmma_kernel(double* out, int iters, double a, double b) {
  double c0[8], c1[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c0[t] = t; c1[t] = -t; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t) dmma884(c0[t], c1[t], a, b);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c0[t] + c1[t];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t fp64_peak_measure(int sms, double* dfma_tflops, double* dmma_tflops, cudaStream_t st) {
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * 1024 * 2 * (size_t)sms);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096, threads = 1024;
  double best[2] = {0.0, 0.0};
  for (int kind = 0; kind < 2; ++kind) {
    const int grid = kind == 0 ? 2 * sms : sms;
    for (int rep = 0; rep < 8; ++rep) {
      cudaEventRecord(e0, st);
      if (kind == 0) peak_dfma_kernel<<<grid, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
      else peak_dmma_kernel<<<grid, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = kind == 0 ? 2.0 * 8 * iters * (double)grid * threads
                                     : 512.0 * 8 * iters * (double)grid * (threads / 32);
      const double tf = flops / (ms * 1e-3) / 1e12;
      if (rep >= 2 && tf > best[kind]) best[kind] = tf;
    }
  }
  *dfma_tflops = best[0];
  *dmma_tflops = best[1];
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError();
}

}  // namespace dab
