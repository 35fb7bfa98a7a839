// FP64 peak microbenchmark for B200 (sm_100a): the roofline denominators that
// MEASURED_PEAKS.json does not carry.  Prints one JSON line.
//   dfma      : independent DFMA chains (the CUDA-core FP64 pipe)
//   dmma884   : mma.sync.aligned.m8n8k4.f64   (native DMMA.8x8x4)
//   dmma16816 : mma.sync.aligned.m16n8k16.f64 (lowered by ptxas to DMMA.8x8x4 sequences)
//   copy      : STREAM-style double2 copy (HBM sanity check against MEASURED_PEAKS.json)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void __launch_bounds__(1024) k_dmma884(double* out, int iters, double a, double b) {
  double c0[TILES], c1[TILES];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c0[t] = t; c1[t] = -t; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[t]), "+d"(c1[t]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c0[t] + c1[t];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void __launch_bounds__(1024) k_dmma16816(double* out, int iters, double a, double b) {
  double c[TILES][4];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c[t][0] = t; c[t][1] = -t; c[t][2] = 1; c[t][3] = 2; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                     "d"(b), "d"(a), "d"(b), "d"(a));
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);

  // DFMA: sweep threads/SM
  for (int threads : {256, 512, 1024}) {
    for (int cps : {1, 2}) {
      if (threads * cps > 2048) continue;
      int grid = sms * cps;
      float ms = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double tf = 2.0 * 8 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
      printf(", \"dfma_t%d_c%d_tflops\": %.3f", threads, cps, tf);
    }
  }
  // DMMA 8x8x4
  for (int threads : {128, 256, 512, 1024}) {
    int grid = sms;
    float ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 512.0 * 8 * iters * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(", \"dmma884_t%d_tflops\": %.3f", threads, tf);
  }
  {
    int grid = sms * 2, threads = 512;
    float ms = time_ms([&] { k_dmma884<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 512.0 * 16 * iters * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(", \"dmma884_t512_c2_x16_tflops\": %.3f", tf);
  }
  // DMMA 16x8x16
  for (int threads : {128, 256, 512}) {
    int grid = sms;
    float ms = time_ms([&] { k_dmma16816<4><<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
    double tf = 4096.0 * 4 * (iters / 4) * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(", \"dmma16816_t%d_tflops\": %.3f", threads, tf);
  }
  // copy
  {
    size_t n = (size_t)1 << 27;  // 128M double2 = 2 GiB per buffer
    double2 *a, *b; CK(cudaMalloc(&a, n * sizeof(double2))); CK(cudaMalloc(&b, n * sizeof(double2)));
    CK(cudaMemset(a, 1, n * sizeof(double2)));
    float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n); }, 5);
    printf(", \"copy_gbs\": %.1f", 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf(", \"clock_khz_max\": %d}\n", clk);
  return 0;
}
