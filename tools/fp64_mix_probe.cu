// Are DFMA (FP64 CUDA-core pipe) and DMMA.8x8x4 (tensor pipe, FP64) separate execution resources on
// B200?  Three kernels of the same shape: all warps DFMA chains, all warps DMMA chains, and half
// the warps each.  If the mixed kernel's combined rate exceeds either pure rate, they overlap.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: DFMA, 1: DMMA, 2: even warps DFMA / odd warps DMMA
__global__ void __launch_bounds__(512) k(double* out, int iters, double a, double b) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = MODE == 1 || (MODE == 2 && (warp & 1));
  double acc[8], c0[4], c1[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { c0[i] = i; c1[i] = -i; }
  if (do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[t]), "+d"(c1[t]) : "d"(a), "d"(b));
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, int sms, double* out) {
  const int iters = 20000, T = 512, grid = sms * 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int r = 0; r < 2; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<grid, T>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double warps = (double)grid * T / 32;
  const double w_mma = MODE == 1 ? warps : (MODE == 2 ? warps / 2 : 0), w_fma = warps - w_mma;
  const double flops = (double)iters * (w_fma * 8 * 32 * 2 + w_mma * 4 * 512);
  printf("%-28s %.3f ms  %.2f TFLOP/s (dfma part %.2f, dmma part %.2f)\n", name, ms, flops / (ms * 1e-3) / 1e12,
         iters * w_fma * 8 * 64 / (ms * 1e-3) / 1e12, iters * w_mma * 4 * 512 / (ms * 1e-3) / 1e12);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 8 * 1024 * 1024);
  run<0>("all warps DFMA", p.multiProcessorCount, out);
  run<1>("all warps DMMA", p.multiProcessorCount, out);
  run<2>("half DFMA, half DMMA", p.multiProcessorCount, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
