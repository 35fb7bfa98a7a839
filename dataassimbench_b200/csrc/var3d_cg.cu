// 3D-Var analysis with a user-supplied background covariance B.
//
// Replaces Var3D._cycle_obsop, dabench/dacycler/_var3d.py:84-110: the "preconditioned with B" system
//     (I + B H^T R^-1 H) xa = xb + B H^T R^-1 y
// solved by conjugate gradients from x0 = xb (jax.scipy.sparse.linalg.cg, tol 1e-5, maxiter 1000: stop when
// ||r||^2 <= tol^2 ||b||^2).  With H the default index gather (masked rows dropped) and R = sd^2 I,
// H^T R^-1 H = diag(w) (w_i = number of valid observations of grid point i / sd^2) and H^T R^-1 y = s (their sum
// / sd^2): the N x n_y and n_y x n_y matrices of the reference never exist, the operator is
//     A p = p + B (w o p),            b = xb + B s,
// i.e. ONE pass over the dense N x N matrix B per iteration -- 8 N^2 bytes, HBM-bound (the only large operand; the
// reference forms B H^T (N x n_y), (B H^T) R^-1, and (...) H (N x N) with three GEMMs before it iterates).
//
// Kernels: `gemv_kernel` (one warp per row of B, 16-byte loads, coalesced along the row, fixed-order shuffle
// reduction), `dot_kernel` + `dot_final_kernel` (two fixed-order stages, no atomics: the iteration is
// bit-reproducible), and two fused vector updates that read alpha / beta from the device scalars, so the only
// host round trip of an iteration is the 8-byte read of ||r||^2 for the stopping test.
#include "common.cuh"
#include "kernels.h"

namespace dab {

namespace {

constexpr int kDotCtas = 148;

// y[i] = add[i] + sum_j B[i][j] * (w ? w[j] * x[j] : x[j])
__global__ void __launch_bounds__(256)
gemv_kernel(const double* __restrict__ B, long long n, const double* __restrict__ w, const double* __restrict__ x,
            const double* __restrict__ add, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    const double* row = B + i * n;
    double s0 = 0.0, s1 = 0.0;
    if ((n & 1) == 0 && (reinterpret_cast<unsigned long long>(row) & 15) == 0) {
      for (long long j = 2 * lane; j < n; j += 64) {
        const double2 b = *reinterpret_cast<const double2*>(row + j);
        const double x0 = w ? w[j] * x[j] : x[j], x1 = w ? w[j + 1] * x[j + 1] : x[j + 1];
        s0 = fma(b.x, x0, s0);
        s1 = fma(b.y, x1, s1);
      }
    } else {
      for (long long j = lane; j < n; j += 32) s0 = fma(row[j], w ? w[j] * x[j] : x[j], s0);
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = (add ? add[i] : 0.0) + s;
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;                                    // valid in thread 0
}

// partial[cta] = sum over the CTA's grid-stride elements of a[i] * b[i]
__global__ void __launch_bounds__(256)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fma(a[i], b[i], acc);
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = sum of the partials (fixed order); optional: out[1] = num[0] / out[0] (alpha = gamma / pAp), or
// out[1] = out[0] / den[0] (beta = gamma_new / gamma)
__global__ void __launch_bounds__(32)
dot_final_kernel(const double* __restrict__ partial, int n_part, double* __restrict__ out, const double* __restrict__ num,
                 const double* __restrict__ den) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n_part; i += 32) t += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) {
    out[0] = t;
    if (num) out[1] = num[0] / t;
    if (den) out[1] = t / den[0];
  }
}

// x += alpha p, r -= alpha Ap, partial[cta] = sum r^2         (alpha = *alpha_dev)
__global__ void __launch_bounds__(256)
cg_update_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ Ap,
                 const double* __restrict__ alpha_dev, long long n, double* __restrict__ partial) {
  __shared__ double red[8];
  const double alpha = *alpha_dev;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, Ap[i], r[i]);
    r[i] = ri;
    acc = fma(ri, ri, acc);
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// p = r + beta p                                              (beta = *beta_dev)
__global__ void __launch_bounds__(256)
cg_direction_kernel(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ beta_dev, long long n) {
  const double beta = *beta_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], r[i]);
}

// r = b - r (r holds A x on entry), p = r, partial[cta] = sum r^2
__global__ void __launch_bounds__(256)
cg_start_kernel(double* __restrict__ r, double* __restrict__ p, const double* __restrict__ b, long long n,
                double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double ri = b[i] - r[i];
    r[i] = ri; p[i] = ri;
    acc = fma(ri, ri, acc);
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// w[i] = den[i] - 1 (the scatter kernel of etkf_contract.cu accumulates 1 + sum 1/sd^2), s[i] = num[i] - xb[i]
__global__ void __launch_bounds__(256)
var3d_ws_kernel(const double* __restrict__ num, const double* __restrict__ den, const double* __restrict__ xb, long long n,
                double* __restrict__ w, double* __restrict__ s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    w[i] = den[i] - 1.0;
    s[i] = num[i] - xb[i];
  }
}

}  // namespace

size_t var3d_cg_workspace_doubles(long long n) { return (size_t)(7 * n) + kDotCtas + 16; }

// xa = CG solution.  ws: var3d_cg_workspace_doubles(n) doubles; loc / val: the valid rows; B [n][n] row-major.
// Synchronises `st` once per iteration (8-byte read of ||r||^2).  *iters_out = iterations made.
cudaError_t var3d_cg_launch(const double* xb, double* xa, long long n, const long long* loc, const double* val,
                            long long n_rows, double w_obs, const double* B, double tol, int maxiter, double* ws,
                            int* bad, int sms, int* iters_out, long* launches, cudaStream_t st) {
  double* wv = ws;            // diag of H^T R^-1 H
  double* sv = wv + n;        // H^T R^-1 y
  double* b = sv + n;
  double* r = b + n;
  double* p = r + n;
  double* Ap = p + n;
  double* num = Ap + n;       // scratch of the scatter
  double* partial = num + n;
  double* sc = partial + kDotCtas;       // [0] gamma, [1] alpha, [2] pAp, [3] alpha, [4] gamma_new, [5] beta, [6] bb
  long long g = (n + 255) / 256;
  if (g > kDotCtas) g = kDotCtas;
  const unsigned gv = (unsigned)g;
  long long gw = (n + 7) / 8;                      // 8 warps (rows) per CTA
  if (gw > (long long)sms * 8) gw = (long long)sms * 8;
  const unsigned gm = (unsigned)gw;
  // ---- w, s from the observation rows (reuses the scatter of the diagonal solver: num = xb + sum, den = 1 + sum)
  cudaError_t e = var3d_scatter_only_launch(xb, num, Ap, n, loc, val, n_rows, w_obs, bad, sms, st);   // Ap: free until the loop
  if (e != cudaSuccess) return e;
  var3d_ws_kernel<<<gv, 256, 0, st>>>(num, Ap, xb, n, wv, sv);
  // ---- b = xb + B s;  ||b||^2;  x = xb;  r = b - A x;  p = r;  gamma = ||r||^2
  gemv_kernel<<<gm, 256, 0, st>>>(B, n, nullptr, sv, xb, b);
  dot_kernel<<<gv, 256, 0, st>>>(b, b, n, partial);
  dot_final_kernel<<<1, 32, 0, st>>>(partial, (int)g, sc + 6, nullptr, nullptr);
  e = cudaMemcpyAsync(xa, xb, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  gemv_kernel<<<gm, 256, 0, st>>>(B, n, wv, xa, xa, r);                  // r := A x
  cg_start_kernel<<<gv, 256, 0, st>>>(r, p, b, n, partial);
  dot_final_kernel<<<1, 32, 0, st>>>(partial, (int)g, sc + 0, nullptr, nullptr);
  if (launches) *launches += 9;
  double host[8];
  e = cudaMemcpyAsync(host, sc, sizeof(double) * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  const double atol2 = tol * tol * host[6];
  double gamma = host[0];
  int it = 0;
  while (gamma > atol2 && it < maxiter) {
    gemv_kernel<<<gm, 256, 0, st>>>(B, n, wv, p, p, Ap);                 // Ap = p + B (w o p)
    dot_kernel<<<gv, 256, 0, st>>>(p, Ap, n, partial);
    dot_final_kernel<<<1, 32, 0, st>>>(partial, (int)g, sc + 2, sc + 0, nullptr);       // pAp, alpha = gamma / pAp
    cg_update_kernel<<<gv, 256, 0, st>>>(xa, r, p, Ap, sc + 3, n, partial);
    dot_final_kernel<<<1, 32, 0, st>>>(partial, (int)g, sc + 4, nullptr, sc + 0);       // gamma_new, beta = gamma_new / gamma
    cg_direction_kernel<<<gv, 256, 0, st>>>(p, r, sc + 5, n);
    e = cudaMemcpyAsync(sc + 0, sc + 4, sizeof(double), cudaMemcpyDeviceToDevice, st);   // gamma := gamma_new
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, sc + 4, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 6;
    gamma = host[0];
    ++it;
  }
  if (iters_out) *iters_out = it;
  return cudaGetLastError();
}

}  // namespace dab
