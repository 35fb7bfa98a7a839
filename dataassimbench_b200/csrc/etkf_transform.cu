// K5 -- apply the ensemble transform: Xa = Xb T with mean update and inflation folded into T.
//
// Replaces the three [N,k]x[k,k] products and the rank-1 mean add of
// dabench/dacycler/_etkf.py:127-138,156-161 with ONE pass over the state:
//   Xa[j][e] = sum_m T[m][j] * Xb[m][src(e)]        (layout [k][N], member-major)
// computed with DMMA.8x8x4 (FP64 tensor core): warp w owns output members 8w..8w+7, the
// T^T fragments stay in registers for the CTA's lifetime, and a k x 64 state tile goes
// through shared memory.
//
// The output is written in the forecast kernel's *extended* layout: every member row carries
// `halo_l` columns before and `halo_r` after its slab.  Halo columns are the transform of the
// periodic neighbours' columns: src(e) = (n0 + e) mod N, read from the owning rank's Xb
// through a peer-mapped pointer (NVLink P2P `ld.global`) -- the wide-halo exchange of the
// sharded forecast is fused into this kernel; with one rank the owner is the rank itself.
#include "common.cuh"
#include "kernels.h"

namespace dab {

constexpr int TC = 64;           // columns per tile
constexpr int XLD = TC + 4;      // smem row stride: (4*m + i) mod 16 distinct -> conflict-free B frags

template <int K8>
__global__ void __launch_bounds__(32 * K8)
transform_kernel(const TransformArgs a) {
  constexpr int KP = 8 * K8;
  extern __shared__ __align__(16) double Xs[];        // [KP][XLD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fr = lane & 3;
  const int k = a.k;

  // A fragments of T^T for this warp's 8 output members: A[jrow][m] = T[m][8*warp + jrow]
  double af[2 * K8];
#pragma unroll
  for (int ks = 0; ks < 2 * K8; ++ks) {
    const int m = 4 * ks + fr, j = 8 * warp + fm;
    af[ks] = (m < k && j < k) ? a.T[m * k + j] : 0.0;
  }

  const long long e_lo = -(long long)a.halo_l;
  const long long n_cols = a.n_loc + a.halo_l + a.halo_r;
  const long long n_tiles = (n_cols + TC - 1) / TC;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long e0 = e_lo + tile * TC;
    __syncthreads();
    // load tile: Xs[m][i] = Xb_owner[m][src col]; a thread keeps one column i (blockDim % TC == 0)
    {
      const int i = tid % TC;
      const long long e = e0 + i;
      const double* sp = nullptr;
      if (e < a.n_loc + a.halo_r) {
        if (e >= 0 && e < a.n_loc) {
          sp = a.src[a.rank] + e;                      // interior: own slab
        } else {                                       // halo: periodic neighbour (peer pointer)
          const long long gp = pmod(a.n0 + e, a.n_global);
          const int owner = (int)(gp / a.n_loc);
          sp = a.src[owner] + (gp - (long long)owner * a.n_loc);
        }
      }
      // 16 independent loads in flight per thread (trip count is a compile-time constant), so
      // L2 / NVLink latency is paid once per tile, not once per row.
      constexpr int ROWS_PER_PASS = (32 * K8) / TC;
      constexpr int PASSES = KP / ROWS_PER_PASS;
      const int m0 = tid / TC;
      double v[PASSES];
#pragma unroll
      for (int it = 0; it < PASSES; ++it) {
        const int m = m0 + it * ROWS_PER_PASS;
        v[it] = (sp != nullptr && m < k) ? sp[(long long)m * a.src_ld] : 0.0;
      }
#pragma unroll
      for (int it = 0; it < PASSES; ++it) Xs[(m0 + it * ROWS_PER_PASS) * XLD + i] = v[it];
    }
    __syncthreads();
    double* drow = a.dst + (long long)(8 * warp + fm) * a.dst_ld + a.dst_base;
    const bool row_ok = (8 * warp + fm) < k;
#pragma unroll 2
    for (int ti = 0; ti < TC / 8; ++ti) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 2 * K8; ++ks) {
        const double b = Xs[(4 * ks + fr) * XLD + 8 * ti + fm];
        dmma884(c0, c1, af[ks], b);
      }
      const long long e = e0 + 8 * ti + 2 * fr;
      if (row_ok) {
        if (e < a.n_loc + a.halo_r) drow[e] = c0;
        if (e + 1 < a.n_loc + a.halo_r) drow[e + 1] = c1;
      }
    }
  }
}

cudaError_t transform_launch(const TransformArgs& a, int sms, cudaStream_t st) {
  const int k8 = (a.k + 7) / 8 <= 2 ? 2 : ((a.k + 7) / 8 <= 4 ? 4 : ((a.k + 7) / 8 <= 8 ? 8 : 16));
  const long long n_cols = a.n_loc + a.halo_l + a.halo_r;
  const long long n_tiles = (n_cols + TC - 1) / TC;
  const size_t smem = (size_t)8 * k8 * XLD * sizeof(double);
  long long g = n_tiles;
  const long long cap = (long long)sms * (k8 <= 8 ? 4 : 2);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
#define DAB_TR_CASE(K)                                                                              \
  {                                                                                                 \
    cudaFuncSetAttribute(transform_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    transform_kernel<K><<<(unsigned)g, 32 * K, smem, st>>>(a);                                      \
  }
  switch (k8) {
    case 2: DAB_TR_CASE(2) break;
    case 4: DAB_TR_CASE(4) break;
    case 8: DAB_TR_CASE(8) break;
    default: DAB_TR_CASE(16) break;
  }
#undef DAB_TR_CASE
  return cudaGetLastError();
}

}  // namespace dab
