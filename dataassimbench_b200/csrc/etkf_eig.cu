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
// Jacobi applied to the columns of G := A: plane rotations from the right until the columns
// are mutually orthogonal; then G = A V = V L, i.e. column c has norm lambda_c and direction
// v_c.  No separate V accumulator is needed (backward stable: V L V^T = A + O(eps |A|)), so the
// whole state is one k x k matrix: 32 KB at k = 64, 128 KB at k = 128 -- one CTA's shared memory.
// Parallel ordering: round-robin tournament, k/2 disjoint pairs per round, one warp per pair;
// the three inner products are warp-shuffle reductions.
#include "common.cuh"
#include "kernels.h"

namespace dab {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stats: [k*k] C row-major, then [k] g.  Tout: [k][k] row-major, T[m][j].
// lam_out (nullable): k eigenvalues of A (unsorted); info_out (nullable): sweeps used.
template <int RPL>   // rows per lane: ceil(k/32)
__global__ void __launch_bounds__(1024)
eig_transform_kernel(const double* __restrict__ stats, int k, double rho, int empty_R,
                     double* __restrict__ Tout, double* __restrict__ lam_out, int* __restrict__ info_out) {
  extern __shared__ __align__(16) double sm[];
  const int kj = (k + 1) & ~1;                 // even number of Jacobi columns (last may be a zero column)
  const int ld = k;                            // column-major, column c at sm + c*ld
  double* G = sm;                              // kj * ld
  double* lam = G + kj * ld;                   // kj
  double* dw = lam + kj;                       // kj   sqrt((k-1)/lambda)
  double* da = dw + kj;                        // kj   1/lambda (0 below the pinv cutoff)
  double* wa = da + kj;                        // k
  double* proj = wa + kj;                      // kj   v_c . g
  double* cmean = proj + kj;                   // k
  __shared__ int s_rot;
  __shared__ double s_red[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const double inv_k = 1.0 / (double)k;

  if (empty_R) {                               // reference zero branch: T = U
    for (int e = tid; e < k * k; e += blockDim.x) Tout[e] = inv_k;
    if (info_out && tid == 0) *info_out = 0;
    return;
  }

  const double shift = (double)(k - 1) / rho;
  for (int e = tid; e < kj * ld; e += blockDim.x) {
    const int c = e / ld, r = e - c * ld;
    double v = 0.0;
    if (c < k) v = stats[r * k + c] + (r == c ? shift : 0.0);
    G[e] = v;
  }
  __syncthreads();

  const int npairs = kj / 2;
  const int m = kj - 1;                        // round-robin modulus
  const double tol = 1.1e-16 * sqrt((double)k) * 4.0;
  int sweeps = 0;
  for (; sweeps < 40; ++sweeps) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    int rotated = 0;
    for (int rnd = 0; rnd < m; ++rnd) {
      for (int pr = warp; pr < npairs; pr += nwarp) {
        int p, q;
        if (pr == 0) { p = m; q = rnd; }
        else { p = (rnd + pr) % m; q = (rnd - pr + m) % m; }
        if (p > q) { const int t_ = p; p = q; q = t_; }
        double* gp = G + p * ld;
        double* gq = G + q * ld;
        double xp[RPL], xq[RPL];
        double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          const int r = lane + 32 * i;
          xp[i] = r < k ? gp[r] : 0.0;
          xq[i] = r < k ? gq[r] : 0.0;
          al = fma(xp[i], xp[i], al);
          be = fma(xq[i], xq[i], be);
          ga = fma(xp[i], xq[i], ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        if (fabs(ga) > tol * sqrt(al * be)) {   // warp-uniform
          rotated = 1;
          const double zeta = (be - al) / (2.0 * ga);
          const double t_ = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
          const double c_ = rsqrt(fma(t_, t_, 1.0));
          const double s_ = c_ * t_;
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = lane + 32 * i;
            if (r < k) {
              gp[r] = c_ * xp[i] - s_ * xq[i];
              gq[r] = s_ * xp[i] + c_ * xq[i];
            }
          }
        }
      }
      __syncthreads();
    }
    if (rotated && lane == 0) s_rot = 1;       // benign race: all writers store 1
    __syncthreads();
    const int any = s_rot;
    __syncthreads();
    if (!any) break;
  }

  // eigenvalues = column norms; normalise columns to eigenvectors
  for (int c = warp; c < k; c += nwarp) {
    double* gc = G + c * ld;
    double ss = 0.0, pg = 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      const double v = r < k ? gc[r] : 0.0;
      ss = fma(v, v, ss);
    }
    ss = warp_sum(ss);
    const double l = sqrt(ss);
    const double il = l > 0.0 ? 1.0 / l : 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + 32 * i;
      if (r < k) {
        const double v = gc[r] * il;
        gc[r] = v;
        pg = fma(v, stats[k * k + r], pg);
      }
    }
    pg = warp_sum(pg);
    if (lane == 0) { lam[c] = l; proj[c] = pg; }
  }
  __syncthreads();
  // pinv cutoff: 1e-15 * largest eigenvalue (rtol semantics of jnp.linalg.pinv)
  double lmax = 0.0;
  for (int c = tid; c < k; c += blockDim.x) lmax = fmax(lmax, lam[c]);
  lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 16));
  lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 8));
  lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 4));
  lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 2));
  lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 1));
  if (lane == 0) s_red[warp] = lmax;
  __syncthreads();
  lmax = 0.0;
  for (int w = 0; w < nwarp; ++w) lmax = fmax(lmax, s_red[w]);
  for (int c = tid; c < k; c += blockDim.x) {
    const double l = lam[c];
    const double inv = l > 1e-15 * lmax ? 1.0 / l : 0.0;
    da[c] = inv;
    dw[c] = sqrt((double)(k - 1) * inv);
    if (lam_out) lam_out[c] = l;
  }
  __syncthreads();
  // wa = V diag(da) V^T g
  for (int r = tid; r < k; r += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < k; ++c) s = fma(G[c * ld + r] * da[c], proj[c], s);
    wa[r] = s;
  }
  // Wa = V diag(dw) V^T  -> Tout (raw), row-major
  for (int e = tid; e < k * k; e += blockDim.x) {
    const int j = e / k, mm = e - j * k;       // consecutive threads: consecutive mm (conflict-free), j broadcast
    double s = 0.0;
#pragma unroll 4
    for (int c = 0; c < k; ++c) s = fma(G[c * ld + mm] * dw[c], G[c * ld + j], s);
    Tout[mm * k + j] = s;
  }
  __syncthreads();
  // column means of Wa and mean of wa
  for (int j = tid; j < k; j += blockDim.x) {
    double s = 0.0;
    for (int mm = 0; mm < k; ++mm) s += Tout[mm * k + j];
    cmean[j] = s * inv_k;
  }
  double wsum = 0.0;
  for (int r = 0; r < k; ++r) wsum += wa[r];   // every thread, same order: identical value
  const double wmean = wsum * inv_k;
  __syncthreads();
  // T[m][j] = 1/k + (wa[m] - mean(wa)) + (Wa[m][j] - colmean_j)
  for (int e = tid; e < k * k; e += blockDim.x) {
    const int mm = e / k, j = e - mm * k;
    Tout[e] = inv_k + (wa[mm] - wmean) + (Tout[e] - cmean[j]);
  }
  if (info_out && tid == 0) *info_out = sweeps;
}

size_t eig_smem_bytes(int k) {
  const int kj = (k + 1) & ~1;
  return ((size_t)kj * k + 7 * (size_t)kj + 8) * sizeof(double);
}

cudaError_t eig_transform_launch(const double* stats, int k, double rho, int empty_R, double* T,
                                 double* lam_out, int* info_out, cudaStream_t st) {
  const size_t smem = eig_smem_bytes(k);
  const int npairs = ((k + 1) & ~1) / 2;
  int warps = npairs < 32 ? npairs : 32;
  if (warps < 1) warps = 1;
  const int threads = 32 * warps;
#define DAB_EIG_CASE(R)                                                                              \
  {                                                                                                  \
    cudaFuncSetAttribute(eig_transform_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    eig_transform_kernel<R><<<1, threads, smem, st>>>(stats, k, rho, empty_R, T, lam_out, info_out);   \
  }
  if (k <= 32) DAB_EIG_CASE(1)
  else if (k <= 64) DAB_EIG_CASE(2)
  else if (k <= 96) DAB_EIG_CASE(3)
  else DAB_EIG_CASE(4)
#undef DAB_EIG_CASE
  return cudaGetLastError();
}

}  // namespace dab
