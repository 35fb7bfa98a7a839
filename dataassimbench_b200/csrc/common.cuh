// Shared helpers for the B200 (sm_100a) ETKF kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace dab {

constexpr int kMaxEns = 128;          // k <= 128 (north star: eigensolver held in one CTA)

// FP64 tensor-core MMA, the native sm_100a shape (SASS: DMMA.8x8x4).
//   A 8x4 row-major : lane l holds A[l/4][l%4]
//   B 4x8 col-major : lane l holds B[l%4][l/4]
//   C/D 8x8         : lane l holds C[l/4][2*(l%4)], C[l/4][2*(l%4)+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ long long pmod(long long a, long long n) {
  long long r = a % n;
  return r < 0 ? r + n : r;
}

}  // namespace dab
