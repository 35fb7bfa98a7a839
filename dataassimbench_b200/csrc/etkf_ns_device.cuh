// K4 for small ensembles (k <= 32) inside ONE CTA: the same deflated, interval-scaled coupled
// Newton-Schulz iteration as etkf_ns.cu (see there for the algebra and the reference lines it
// replaces, dabench/dacycler/_etkf.py:142-161), with the three 32 x 32 matrices in shared memory
// and the products on the FP64 tensor core.  Used by the batched small-problem kernel
// (etkf_batched.cu: one CTA per experiment) and by the standalone K4 launch for k <= 32.
// The Jacobi eigensolver it replaces there needs ~150 barrier-separated rotation rounds per
// analysis (k = 20); this needs 5-6 steps of three barrier-separated products.
#pragma once
#include "common.cuh"

namespace dab {

constexpr double kNsSmallMaxCond = 1e7;

// padded size KP = 16, 24 or 32 (identity on the padding): the work goes with KP^3
__host__ __device__ inline int ns_small_kp(int k) { return k <= 16 ? 16 : (k <= 24 ? 24 : 32); }
__host__ __device__ inline int ns_small_smem_doubles() { return 3 * 32 * 36 + 6 * 32 + 8; }   // the KP = 32 case

// stats: [k*k] C row-major then [k] g (global or shared).  Tout: [k][k] row-major (global or shared).
// Returns false (uniformly over the CTA, nothing written) when the statistics are not finite or the
// bound on cond(A') exceeds kNsSmallMaxCond: the caller then runs the Jacobi eigensolver.
// Must be called by every thread of the CTA; blockDim.x == 256 (returns false otherwise).
template <int KP>
static __device__ __noinline__ bool ns_small_body_kp(const double* stats, int k, double rho, double* Tout,
                                                     int* info_out, double* sm) {
  constexpr int LD = KP + 4;                   // row stride: DMMA fragments conflict-free both ways
  double* Y = sm;                              // Y_k          [KP][LD]
  double* Z = Y + KP * LD;                     // Z_k
  double* Tm = Z + KP * LD;                    // T_k = c1 I + c2 Z Y
  double* vec = Tm + KP * LD;                  // 6*KP + 8 scratch
  double* r_sq = vec, *r_abs = vec + KP, *r_g = vec + 2 * KP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
  const int fm = lane >> 2, fr = lane & 3;
  const double shift = (double)(k - 1) / rho;
  if (nwarp != 8) return false;                // 8 warps x 2 tiles: the mapping below is fixed to it

  // ---- A' = A + delta U (deflation of the known eigenvector 1), bounds on its spectrum
  double tr = 0.0;
  for (int i = lane; i < k; i += 32) tr += stats[i * k + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  double delta_k = k > 1 ? tr / ((double)(k - 1) * (double)k) : 0.0;
  if (!(delta_k >= 0.0)) delta_k = 0.0;
  for (int i = warp; i < KP; i += nwarp) {     // one warp per row, one column per lane
    const int j = lane;                        // lanes >= KP idle (KP <= 32)
    double a = (i == j) ? 1.0 : 0.0;           // identity on the padding (scaled below)
    const bool in = i < k && j < k;
    if (in) a = stats[i * k + j] + delta_k + (i == j ? shift : 0.0);
    double sq = in ? a * a : 0.0, ab = in ? fabs(a) : 0.0, di = (in && i == j) ? a : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      ab += __shfl_xor_sync(0xffffffffu, ab, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
    }
    if (j < KP) {
      Y[i * LD + j] = a;                       // unscaled for now
      Z[i * LD + j] = (i == j) ? 1.0 : 0.0;
    }
    if (lane == 0) { r_sq[i] = sq; r_abs[i] = ab; r_g[i] = i < k ? 2.0 * di - ab : 1e300; }
  }
  __syncthreads();
  double fro2 = lane < KP ? r_sq[lane] : 0.0, rmax = lane < KP ? r_abs[lane] : 0.0, gmin = lane < KP ? r_g[lane] : 1e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fro2 += __shfl_xor_sync(0xffffffffu, fro2, o);
    const double a = __shfl_xor_sync(0xffffffffu, rmax, o);
    rmax = a > rmax ? a : rmax;
    const double g_ = __shfl_xor_sync(0xffffffffu, gmin, o);
    gmin = g_ < gmin ? g_ : gmin;
  }
  const double fro = sqrt(fro2);
  const double s = fro < rmax ? fro : rmax;
  const double lower = gmin > shift ? gmin : shift;
  const bool ok = (fro2 < 1e300) && (s > 0.0) && (s <= kNsSmallMaxCond * lower);
  if (!ok) { __syncthreads(); return false; }
  const double inv_s = 1.0 / s;
  for (int e = tid; e < KP * KP; e += nthr) {
    const int i = e / KP, j = e - i * KP;
    if (i < k && j < k) Y[i * LD + j] *= inv_s;
  }
  __syncthreads();

  // ---- coupled iteration, full matrices: M = Z Y, T = c1 I + c2 M, Y <- Y T, Z <- T Z
  double l = sqrt(lower * inv_s) * (1.0 - 1e-6);
  int iters = 0;
  constexpr int NT = KP / 8;                   // 8 x 8 tiles per side
  constexpr int TPW = (NT * NT + 7) / 8;       // tiles per warp (8 warps)
  for (; iters < 40; ++iters) {
    const double al2 = 3.0 / (1.0 + l + l * l), al = sqrt(al2);
    const double c1 = 1.5 * al, c2 = -0.5 * al * al2;
    double m0[TPW], m1[TPW];                                  // T tiles (ti, tj) of this warp: all loads, then the stores
#pragma unroll
    for (int c = 0; c < TPW; ++c) {
      const int t = warp + 8 * c, ti = t / NT, tj = t - ti * NT;
      m0[c] = 0.0; m1[c] = 0.0;
      if (t < NT * NT) {                                      // warp-uniform
#pragma unroll
        for (int ks = 0; ks < KP / 4; ++ks)
          dmma884(m0[c], m1[c], Z[(8 * ti + fm) * LD + 4 * ks + fr], Y[(4 * ks + fr) * LD + 8 * tj + fm]);
      }
    }
#pragma unroll
    for (int c = 0; c < TPW; ++c) {
      const int t = warp + 8 * c, ti = t / NT, tj = t - ti * NT;
      if (t >= NT * NT) continue;
      const int row = 8 * ti + fm, col = 8 * tj + 2 * fr;
      Tm[row * LD + col] = fma(c2, m0[c], row == col ? c1 : 0.0);
      Tm[row * LD + col + 1] = fma(c2, m1[c], row == col + 1 ? c1 : 0.0);
    }
    __syncthreads();
    // both products of a tile position in one pass; results held until every warp has read Y and Z
    double y0[TPW], y1[TPW], z0[TPW], z1[TPW];
#pragma unroll
    for (int c = 0; c < TPW; ++c) {
      const int t = warp + 8 * c, ti = t / NT, tj = t - ti * NT;
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      if (t < NT * NT) {
#pragma unroll
        for (int ks = 0; ks < KP / 4; ++ks) {
          dmma884(a0, a1, Y[(8 * ti + fm) * LD + 4 * ks + fr], Tm[(4 * ks + fr) * LD + 8 * tj + fm]);   // Y T
          dmma884(b0, b1, Tm[(8 * ti + fm) * LD + 4 * ks + fr], Z[(4 * ks + fr) * LD + 8 * tj + fm]);   // T Z
        }
      }
      y0[c] = a0; y1[c] = a1; z0[c] = b0; z1[c] = b1;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < TPW; ++c) {
      const int t = warp + 8 * c, ti = t / NT, tj = t - ti * NT;
      if (t >= NT * NT) continue;
      const int row = 8 * ti + fm, col = 8 * tj + 2 * fr;
      Y[row * LD + col] = y0[c]; Y[row * LD + col + 1] = y1[c];
      Z[row * LD + col] = z0[c]; Z[row * LD + col + 1] = z1[c];
    }
    __syncthreads();
    l = 0.5 * al * l * (3.0 - al2 * l * l);
    if (1.0 - l < 1e-12) { ++iters; break; }
  }

  // ---- T = U + (I-U)(wa 1^T + Wa),  Wa = sqrt((k-1)/s) Z,  wa = Z (Z g) / s
  double* gs = vec, *us = vec + KP, *cs = vec + 2 * KP, *was = vec + 3 * KP;
  for (int j = tid; j < KP; j += nthr) gs[j] = j < k ? stats[k * k + j] : 0.0;
  __syncthreads();
  for (int i = warp; i < KP; i += nwarp) {     // u_i = Z[i,:] g and cs_i = sum_{r<k} Z[r][i]
    double u = lane < KP ? Z[i * LD + lane] * gs[lane] : 0.0;
    double c = lane < k ? Z[lane * LD + i] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { u += __shfl_xor_sync(0xffffffffu, u, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { us[i] = u; cs[i] = c; }
  }
  __syncthreads();
  const double inv_k = 1.0 / (double)k;
  for (int i = warp; i < KP; i += nwarp) {
    double w = lane < KP ? Z[i * LD + lane] * us[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) was[i] = w * inv_s;
  }
  __syncthreads();
  double wmean = lane < k ? was[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmean += __shfl_xor_sync(0xffffffffu, wmean, o);
  wmean *= inv_k;
  const double cw = sqrt((double)(k - 1) * inv_s);
  for (int e = tid; e < k * k; e += nthr) {
    const int i = e / k, j = e - i * k;
    Tout[e] = inv_k + (was[i] - wmean) + cw * (Z[i * LD + j] - cs[j] * inv_k);
  }
  if (info_out && tid == 0) *info_out = iters;
  __syncthreads();
  return true;
}

static __device__ __forceinline__ bool ns_small_body(const double* stats, int k, double rho, double* Tout,
                                                     int* info_out, double* sm) {
  if (k <= 16) return ns_small_body_kp<16>(stats, k, rho, Tout, info_out, sm);
  if (k <= 24) return ns_small_body_kp<24>(stats, k, rho, Tout, info_out, sm);
  return ns_small_body_kp<32>(stats, k, rho, Tout, info_out, sm);
}

}  // namespace dab
