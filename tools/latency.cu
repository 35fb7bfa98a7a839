// Dependent-issue latencies on B200 (cycles per op in a single-warp dependent chain).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  __shared__ double sm[64];
  sm[threadIdx.x] = x; __syncthreads();
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, b, a);
    if (OP == 1) x = x + b;
    if (OP == 2) x = x * b;
    if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == 4) x = rsqrt(x) + 1.0;
    if (OP == 5) x = 1.0 / x + 1.0;
    if (OP == 6) x = sqrt(x) + 1.0;
    if (OP == 7) { float f = __frsqrt_rn((float)x); x = (double)f + 1.0; }
    if (OP == 8) { int j = ((int)x) & 31; x = sm[j] + 1.0; }
    if (OP == 9) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(a) : "d"(b), "d"(b)); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 512); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "dadd", "dmul", "shfl64", "rsqrt+add", "div+add", "sqrt+add", "f32rsqrt+cvt+add", "lds+cvt+add", "dmma"};
  printf("{");
  for (int op = 0; op < 10; ++op) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 1: k<1><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 2: k<2><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 3: k<3><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 4: k<4><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 5: k<5><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 6: k<6><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 7: k<7><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 8: k<8><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
        case 9: k<9><<<1, 32>>>(out, cyc, 1.0000001, 0.999999); break;
      }
      cudaDeviceSynchronize();
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("%s\"%s\": %.1f", op ? ", " : "", names[op], (double)c / N);
  }
  printf("}\n");
  return 0;
}
