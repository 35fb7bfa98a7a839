// K4 -- k x k symmetric eigendecomposition in one CTA's shared memory, and the transform T.
//
// Replaces (dabench/dacycler/_etkf.py):
//   Pa_ens = pinv((k-1)/rho I + Yb_pert.T Rinv Yb_pert, rtol=1e-15)      :144-146
//   Wa     = sqrtm((k-1) Pa_ens).real                                     :147-148
//   empty-R branch (Pa = Wa = 0)                                          :149-152
//   wa     = Pa_ens Yb_pert.T Rinv (y - yb_bar)                           :154
//   Xa = Xb_pert Wa + (xb_bar + Xb_pert wa) 1^T                           :156-161
// by the single matrix  T = U + (I-U)(wa 1^T + Wa),  U = 11^T/k,  so that Xa = Xb T  (K5).
//
// Method: A = (k-1)/rho I + C is symmetric positive definite.  One-sided (Hestenes) cyclic
// Jacobi on the columns of G := A: plane rotations from the right until the columns are
// mutually orthogonal; then G = A V = V L, i.e. column c has norm lambda_c and direction v_c.
// No separate V accumulator is needed (backward stable: V L V^T = A + O(eps |A|)), so the whole
// state is one k x k matrix: 36 KB at k = 64, 139 KB at k = 128 -- one CTA's shared memory.
//
// Mapping (latency- and issue-bound, so every instruction counts):
//   * round-robin tournament, k/2 disjoint pairs per round, one __syncthreads per round;
//   * a pair is owned by LPP lanes holding 8 rows each (LPP = 8 at k = 64: four pairs per warp),
//     so the three inner products cost 3*log2(LPP) shuffles and one 8-deep FMA chain;
//   * thresholds are tested as gamma^2 > tol^2 alpha beta (no sqrt/div), and the rotation is
//     r = rsqrt(tau^2 + 4 gamma^2), x = (1 + |tau| r)/2, c = x rsqrt(x), s = sgn(tau) gamma r rsqrt(x)
//     -- two rsqrt instead of two divisions and two square roots;
//   * sweeps stop when nothing rotated, or one sweep early once every |gamma|/sqrt(alpha beta)
//     seen in a sweep was below 1e-8 (quadratic convergence finishes the job within that sweep);
//   * Wa = V diag(sqrt((k-1)/lambda)) V^T runs on the FP64 tensor core (DMMA.8x8x4).
#include <cstdlib>

#include "common.cuh"
#include "etkf_eig_device.cuh"
#include "etkf_ns_device.cuh"
#include "kernels.h"

namespace dab {

// k <= 32 without an eigenvalue request: in-CTA Newton-Schulz (etkf_ns_device.cuh), with the Jacobi
// body as its fallback in the same kernel.  256 threads.
template <int LPP>
__global__ void __launch_bounds__(256)
ns_small_transform_kernel(const double* __restrict__ stats, int k, double rho, double* __restrict__ Tout,
                          int* __restrict__ info_out) {
  extern __shared__ __align__(16) double sm[];
  if (!ns_small_body(stats, k, rho, Tout, info_out, sm))
    eig_small_body<LPP>(stats, k, rho, 0, Tout, nullptr, info_out, sm);
}

template <int LPP>
__global__ void __launch_bounds__(1024)
eig_transform_kernel(const double* __restrict__ stats, int k, double rho, int empty_R,
                     double* __restrict__ Tout, double* __restrict__ lam_out, int* __restrict__ info_out) {
  extern __shared__ __align__(16) double sm[];
  eig_small_body<LPP>(stats, k, rho, empty_R, Tout, lam_out, info_out, sm);
}

template <int RPL>
__global__ void __launch_bounds__(512)
eig_block_kernel(const double* __restrict__ stats, int k, double rho,
                 double* __restrict__ Tout, double* __restrict__ lam_out, int* __restrict__ info_out, int skew,
                 const int* __restrict__ run_flag) {
  // enqueued behind the Newton-Schulz kernel (etkf_ns.cu) as its fallback: runs only if asked to
  if (run_flag != nullptr && *run_flag == 0) return;
  extern __shared__ __align__(16) double sm[];
  eig_block_body<RPL>(stats, k, (k + 7) & ~7, rho, Tout, lam_out, info_out, skew, sm);
}

static int eig_lpp(int k) { return eig_small_lpp(k); }

size_t eig_smem_bytes(int k) {
  if (k > 32) {
    const int rows = k <= 64 ? 64 : 128;
    const int ncol = (k + 7) & ~7;
    return ((size_t)ncol * (rows + 8) + 7 * (size_t)ncol + 8) * sizeof(double);
  }
  const int lpp = eig_lpp(k);
  const int ld = 8 * lpp + 8;
  const int kj = (k + 1) & ~1, k8 = (k + 7) >> 3;
  const int ncol = kj > 8 * k8 ? kj : 8 * k8;
  return ((size_t)ncol * ld + 6 * (size_t)ncol + 8) * sizeof(double);
}

cudaError_t eig_transform_launch(const double* stats, int k, double rho, int empty_R, double* T,
                                 double* lam_out, int* info_out, double* ns_ws, int method,
                                 cudaStream_t st) {
  const size_t smem = eig_smem_bytes(k);
  if (k > 32 && !empty_R) {
    // Newton-Schulz on a CTA cluster unless eigenvalues are wanted; for k > 64 Jacobi stays enqueued
    // behind it as the fallback for ill-conditioned or non-finite statistics (one flag read when not
    // needed); for k <= 64 the fallback runs inside the Newton-Schulz kernel.
    const int* run_flag = nullptr;
    if (method != kEigJacobi && lam_out == nullptr && ns_ws != nullptr) {
      cudaError_t e = ns_transform_launch(stats, k, rho, ns_ws, T, info_out, st);
      if (e != cudaSuccess) return e;
      if (ns_fallback_in_kernel(k)) return cudaGetLastError();    // the fallback for k <= 64 lives inside that kernel
      run_flag = reinterpret_cast<const int*>(ns_ws);
    }
    const int threads = 32 * (((k + 7) & ~7) / 8);
    int skew = 150;          // cycles; measured optimum 100-200 on B200 (4 % faster than 0)
    if (const char* ev = getenv("DAB_EIG_SKEW")) skew = atoi(ev);
    if (k <= 64) {
      cudaFuncSetAttribute(eig_block_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      eig_block_kernel<8><<<1, threads, smem, st>>>(stats, k, rho, T, lam_out, info_out, skew, run_flag);
    } else {
      cudaFuncSetAttribute(eig_block_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      eig_block_kernel<16><<<1, threads, smem, st>>>(stats, k, rho, T, lam_out, info_out, skew, run_flag);
    }
    return cudaGetLastError();
  }
  const int lpp = eig_lpp(k);
  if (k <= 32 && !empty_R && method != kEigJacobi && lam_out == nullptr) {
    const size_t ns_smem = sizeof(double) * (size_t)(ns_small_smem_doubles() > eig_small_smem_doubles(k)
                                                         ? ns_small_smem_doubles() : eig_small_smem_doubles(k));
#define DAB_NS_SMALL_CASE(L)                                                                              \
  {                                                                                                       \
    cudaFuncSetAttribute(ns_small_transform_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ns_smem); \
    ns_small_transform_kernel<L><<<1, 256, ns_smem, st>>>(stats, k, rho, T, info_out);                     \
  }
    switch (lpp) {
      case 1: DAB_NS_SMALL_CASE(1) break;
      case 2: DAB_NS_SMALL_CASE(2) break;
      default: DAB_NS_SMALL_CASE(4) break;
    }
#undef DAB_NS_SMALL_CASE
    return cudaGetLastError();
  }
  const int npairs = ((k + 1) & ~1) / 2;
  const int ppw = 32 / lpp;
  int warps = (npairs + ppw - 1) / ppw;
  if (warps < 1) warps = 1;
  if (warps > 32) warps = 32;
  const int threads = 32 * warps;
#define DAB_EIG_CASE(L)                                                                              \
  {                                                                                                  \
    cudaFuncSetAttribute(eig_transform_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    eig_transform_kernel<L><<<1, threads, smem, st>>>(stats, k, rho, empty_R, T, lam_out, info_out);   \
  }
  switch (lpp) {
    case 1: DAB_EIG_CASE(1) break;
    case 2: DAB_EIG_CASE(2) break;
    case 4: DAB_EIG_CASE(4) break;
    case 8: DAB_EIG_CASE(8) break;
    default: DAB_EIG_CASE(16) break;
  }
#undef DAB_EIG_CASE
  return cudaGetLastError();
}

}  // namespace dab
