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
#include "kernels.h"

namespace dab {

template <int LPP>
__global__ void __launch_bounds__(1024)
eig_transform_kernel(const double* __restrict__ stats, int k, double rho, int empty_R,
                     double* __restrict__ Tout, double* __restrict__ lam_out, int* __restrict__ info_out) {
  extern __shared__ __align__(16) double sm[];
  eig_small_body<LPP>(stats, k, rho, empty_R, Tout, lam_out, info_out, sm);
}

// ---- warp-local block Jacobi (32 < k <= 128) ------------------------------------------------
// Columns are grouped in blocks of 4; a warp keeps a PAIR of blocks (8 columns, 8 lanes x RPL rows
// per column pair) in registers for a whole block-round and rotates the 16 cross pairs in four
// sub-rounds, passing the second block's columns from lane group to lane group with shuffles.
// Shared memory is touched, and the CTA synchronised, once per block-round (15 per sweep at
// k = 64) instead of once per rotation round (63); the three rounds of within-block pairs go
// through shared memory.  Column norms ride along with the columns (updated by the rotation,
// recomputed at every load), so a sub-round needs one inner product instead of three.
template <int RPL>
__global__ void __launch_bounds__(512)
eig_block_kernel(const double* __restrict__ stats, int k, double rho,
                 double* __restrict__ Tout, double* __restrict__ lam_out, int* __restrict__ info_out, int skew,
                 const int* __restrict__ run_flag) {
  // enqueued behind the Newton-Schulz kernel (etkf_ns.cu) as its fallback: runs only if asked to
  if (run_flag != nullptr && *run_flag == 0) return;
  constexpr int LPP = 8;
  constexpr int ROWS = RPL * LPP;
  constexpr int LD = ROWS + 8;
  extern __shared__ __align__(16) double sm[];
  const int ncol = (k + 7) & ~7;               // 4-column blocks, even number of blocks
  const int nb = ncol >> 2;
  double* G = sm;
  double* aux = G + ncol * LD;
  double* nrm = aux + 6 * ncol;                // squared column norms, carried with the columns
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, grp = lane >> 3;

  const double shift = (double)(k - 1) / rho;
  for (int e = tid; e < ncol * LD; e += blockDim.x) {
    const int c = e / LD, r = e - c * LD;
    double v = 0.0;
    if (c < k && r < k) v = stats[r * k + c] + (r == c ? shift : 0.0);
    G[e] = v;
  }
  __syncthreads();

  const double tol = 1.1e-16 * sqrt((double)k) * 4.0;
  const double tol2 = tol * tol;
  const double quad2 = 1e-16;
  const int m = nb - 1;
  int sweeps = 0;
#ifdef DAB_EIG_TIMING
  long long tA = 0, tLoad = 0, tSub = 0, tStore = 0, tFlags = 0, tick_ = clock64();
#endif
  for (; sweeps < 40; ++sweeps) {
    int rotated = 0, big = 0;
    DAB_TICK(tFlags);
    // ---- phase A: the 6 pairs inside every block, 3 rounds through shared memory
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      const int gid = warp * 4 + grp;           // 2 lane groups per block
      const int blk = gid >> 1, wh = gid & 1;
      int p, q;
      if (t == 0) { p = 2 * wh; q = 2 * wh + 1; }
      else if (t == 1) { p = wh; q = wh + 2; }
      else { p = wh; q = 3 - wh; }
      double* gp = G + (4 * blk + p) * LD + sub;
      double* gq = G + (4 * blk + q) * LD + sub;
      double xp[RPL], xq[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) { xp[i] = gp[LPP * i]; xq[i] = gq[LPP * i]; }
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        al = fma(xp[i], xp[i], al); be = fma(xq[i], xq[i], be); ga = fma(xp[i], xq[i], ga);
      }
      al = group_sum<LPP>(al); be = group_sum<LPP>(be); ga = group_sum<LPP>(ga);
      const double g2 = ga * ga, ab = al * be;
      if (g2 > tol2 * ab) {
        rotated = 1;
        if (g2 > quad2 * ab) big = 1;
        const double tau = be - al;
        const double q_ = fma(tau, tau, 4.0 * g2);
        const double r = rsqrt(q_);
        const double x = fma(0.5 * fabs(tau), r, 0.5);
        const double ic = rsqrt(x);
        const double c_ = x * ic;
        const double s_ = copysign(ga * r * ic, tau >= 0.0 ? ga : -ga);
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          gp[LPP * i] = fma(c_, xp[i], -s_ * xq[i]);
          gq[LPP * i] = fma(s_, xp[i], c_ * xq[i]);
        }
        // the rotation diagonalises the 2x2 Gram: norms' = (al + be -/+ sgn(tau) h) / 2, h = sqrt(q)
        const double hh = copysign(q_ * r, tau);
        const double sum = al + be;
        al = 0.5 * (sum - hh); be = 0.5 * (sum + hh);
      }
      if (t == 2 && sub == 0) { nrm[4 * blk + p] = al; nrm[4 * blk + q] = be; }
      __syncthreads();
    }
    DAB_TICK(tA);
    // ---- phase B: block round-robin, 16 cross pairs per block pair in registers
#pragma unroll 1
    for (int R = 0; R < m; ++R) {
      int I, J;
      if (warp == 0) { I = m; J = R; }
      else {
        I = R + warp; if (I >= m) I -= m;
        J = R - warp; if (J < 0) J += m;
      }
      if (skew > 0 && (warp & 4)) {              // de-phase the two warps that share an SMSP
        const long long t_end = clock64() + skew;
        while (clock64() < t_end) { }
      }
      double* ga_ptr = G + (4 * I + grp) * LD + sub;
      double xa[RPL], xb[RPL];
      double na = nrm[4 * I + grp], nbn = nrm[4 * J + grp];
      {
        const double* gb_ptr = G + (4 * J + grp) * LD + sub;
#pragma unroll
        for (int i = 0; i < RPL; ++i) { xa[i] = ga_ptr[LPP * i]; xb[i] = gb_ptr[LPP * i]; }
      }
#ifdef DAB_EIG_TIMING
      if (xa[0] + xb[0] + na == -1.2345e300) tLoad -= 1;   // make the loads complete before the tick
#endif
      DAB_TICK(tLoad);
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        double g0 = xa[0] * xb[0], g1 = xa[1] * xb[1];
#pragma unroll
        for (int i = 2; i < RPL; i += 2) { g0 = fma(xa[i], xb[i], g0); g1 = fma(xa[i + 1], xb[i + 1], g1); }
        const double ga = group_sum<LPP>(g0 + g1);
        const double g2 = ga * ga, ab = na * nbn;
        const bool rot = g2 > tol2 * ab;
        const int src = (lane + 8) & 31;         // the b columns move to the previous lane group
        if (__any_sync(0xffffffffu, rot)) {      // warp-uniform: lets the shuffles overlap the rotation
          double c_ = 1.0, s_ = 0.0;
          if (rot) {
            rotated = 1;
            if (g2 > quad2 * ab) big = 1;
            const double tau = nbn - na;
            const double q_ = fma(tau, tau, 4.0 * g2);
            const double r = rsqrt(q_);
            const double x = fma(0.5 * fabs(tau), r, 0.5);
            const double ic = rsqrt(x);
            c_ = x * ic;
            s_ = copysign(ga * r * ic, tau >= 0.0 ? ga : -ga);
            // norms' = (na + nb -/+ sgn(tau) h) / 2 with h = sqrt(q) = q r: off the c/s chain
            const double hh = copysign(q_ * r, tau);
            const double sum = na + nbn;
            na = 0.5 * (sum - hh); nbn = 0.5 * (sum + hh);
          }
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const double a_ = xa[i], b_ = xb[i];
            xa[i] = fma(c_, a_, -s_ * b_);
            const double nb_ = fma(s_, a_, c_ * b_);
            xb[i] = (sr < 3) ? __shfl_sync(0xffffffffu, nb_, src) : nb_;
          }
          if (sr < 3) nbn = __shfl_sync(0xffffffffu, nbn, src);
        } else if (sr < 3) {
#pragma unroll
          for (int i = 0; i < RPL; ++i) xb[i] = __shfl_sync(0xffffffffu, xb[i], src);
          nbn = __shfl_sync(0xffffffffu, nbn, src);
        }
      }
#ifdef DAB_EIG_TIMING
      if (xa[0] + xb[0] == -1.2345e300) tSub -= 1;
#endif
      DAB_TICK(tSub);
      double* gb_out = G + (4 * J + ((grp + 3) & 3)) * LD + sub;
#pragma unroll
      for (int i = 0; i < RPL; ++i) { ga_ptr[LPP * i] = xa[i]; gb_out[LPP * i] = xb[i]; }
      if (sub == 0) { nrm[4 * I + grp] = na; nrm[4 * J + ((grp + 3) & 3)] = nbn; }
      __syncthreads();
      DAB_TICK(tStore);
    }
    const int any_rot = __syncthreads_or(rotated);
    const int any_big = __syncthreads_or(big);
    if (!any_rot) break;
    if (!any_big) { ++sweeps; break; }
  }
  DAB_TICK(tFlags);
  build_transform<LD, ROWS>(G, aux, ncol, stats, k, Tout, lam_out, info_out, sweeps);
#ifdef DAB_EIG_TIMING
  long long tBuild = 0;
  DAB_TICK(tBuild);
  if (tid == 0 && lam_out) {
    lam_out[0] = (double)tA; lam_out[1] = (double)tLoad; lam_out[2] = (double)tSub;
    lam_out[3] = (double)tStore; lam_out[4] = (double)tFlags; lam_out[5] = (double)tBuild;
  }
#endif
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
    // Newton-Schulz on a CTA cluster unless eigenvalues are wanted; Jacobi stays enqueued behind it
    // as the fallback for ill-conditioned or non-finite statistics (one flag read when not needed).
    const int* run_flag = nullptr;
    if (method != kEigJacobi && lam_out == nullptr && ns_ws != nullptr) {
      cudaError_t e = ns_transform_launch(stats, k, rho, ns_ws, T, info_out, st);
      if (e != cudaSuccess) return e;
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
