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

  // Balanced split: the extended row is cut into 8-column DMMA tiles and every CTA takes the same
  // number of them (+-1) as ONE contiguous range, walked in chunks of up to TC columns -- no
  // last-round stragglers (1028 64-column tiles on 592 CTAs used to leave 26 % of the second round idle).
  const long long e_lo = -(long long)a.halo_l;
  const long long e_hi = a.n_loc + a.halo_r;
  const long long n8 = (e_hi - e_lo + 7) / 8;
  const long long t_begin = n8 * blockIdx.x / gridDim.x, t_end = n8 * (blockIdx.x + 1) / gridDim.x;
  // a thread keeps one column i of a chunk (blockDim % TC == 0) and PASSES rows of it
  constexpr int ROWS_PER_PASS = (32 * K8) / TC;
  constexpr int PASSES = KP / ROWS_PER_PASS;
  const int i = tid % TC, m0 = tid / TC;
  double v[PASSES];
  // chunk [t0, t0 + 8) -> registers: Xb_owner[m][src col]; all PASSES loads are independent, so
  // L2 / NVLink latency is paid once per chunk, not once per row
  auto fetch = [&](long long t0) {
    const int tiles_here = (int)(t_end - t0 < TC / 8 ? t_end - t0 : TC / 8);
    const long long e = e_lo + t0 * 8 + i;
    const double* sp = nullptr;
    if (t0 < t_end && i < 8 * tiles_here && e < e_hi) {
      if (e >= 0 && e < a.n_loc) {
        sp = a.src[a.rank] + e;                      // interior: own slab
      } else {                                       // halo: periodic neighbour (peer pointer)
        const long long gp = pmod(a.n0 + e, a.n_global);
        const int owner = (int)(gp / a.n_loc);
        sp = a.src[owner] + (gp - (long long)owner * a.n_loc);
      }
    }
#pragma unroll
    for (int it = 0; it < PASSES; ++it) {
      const int m = m0 + it * ROWS_PER_PASS;
      v[it] = (sp != nullptr && m < k) ? sp[(long long)m * a.src_ld] : 0.0;
    }
  };
  // The background ensemble was finished two kernels ago (see pdl_wait in common.cuh: this CTA can
  // only be resident once the kernel that made T has passed its own wait), so the first chunk is
  // fetched while T is still being computed; T itself is read after the wait.
  fetch(t_begin);
  pdl_wait();
  pdl_launch_dependents();
  // A fragments of T^T for this warp's 8 output members: A[jrow][m] = T[m][8*warp + jrow]
  double af[2 * K8];
#pragma unroll
  for (int ks = 0; ks < 2 * K8; ++ks) {
    const int m = 4 * ks + fr, j = 8 * warp + fm;
    af[ks] = (m < k && j < k) ? a.T[m * k + j] : 0.0;
  }
  for (long long t0 = t_begin; t0 < t_end; t0 += TC / 8) {
    const int tiles_here = (int)(t_end - t0 < TC / 8 ? t_end - t0 : TC / 8);
    const long long e0 = e_lo + t0 * 8;
    __syncthreads();                                 // every warp is done with the previous chunk
#pragma unroll
    for (int it = 0; it < PASSES; ++it) Xs[(m0 + it * ROWS_PER_PASS) * XLD + i] = v[it];
    __syncthreads();
    fetch(t0 + TC / 8);                              // next chunk in flight behind this chunk's DMMAs
    double* drow = a.dst + (long long)(8 * warp + fm) * a.dst_ld + a.dst_base;
    const bool row_ok = (8 * warp + fm) < k;
    const bool vec_ok = ((a.dst_ld | a.dst_base) & 1) == 0 && (reinterpret_cast<unsigned long long>(a.dst) & 15) == 0;
#pragma unroll 2
    for (int ti = 0; ti < tiles_here; ++ti) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 2 * K8; ++ks) {
        const double b = Xs[(4 * ks + fr) * XLD + 8 * ti + fm];
        dmma884(c0, c1, af[ks], b);
      }
      const long long e = e0 + 8 * ti + 2 * fr;
      if (row_ok) {
        if (vec_ok && e + 1 < e_hi && ((e - e_lo) & 1) == 0 && (a.halo_l & 1) == 0) {
          *reinterpret_cast<double2*>(drow + e) = make_double2(c0, c1);
        } else {
          if (e < e_hi) drow[e] = c0;
          if (e + 1 < e_hi) drow[e + 1] = c1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// On-device score of an analysis (the step after the path, SURVEY.md 8f: dabench.metrics rmse/mse
// of the ensemble mean, dabench/metrics/_deterministic.py:43-72, evaluated without copying the
// analysis to the host): sse = sum over this rank's columns of (mean_j Xa[j][e] - truth[e])^2.
// Two fixed-order stages (per-CTA partials, then one warp): bit-reproducible, no atomics.
constexpr int kSseCtas = 296;

__global__ void __launch_bounds__(256)
mean_sqerr_kernel(const double* __restrict__ xa, long long ld, int k, long long n, const double* __restrict__ truth,
                  double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  const double inv_k = 1.0 / (double)k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    double m = 0.0;
    for (int j = 0; j < k; ++j) m += xa[(long long)j * ld + e];
    const double d = m * inv_k - truth[e];
    acc = fma(d, d, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32)
sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) *out = t;
}

cudaError_t mean_sqerr_launch(const double* xa, long long ld, int k, long long n, const double* truth,
                              double* partial /* kSseCtas doubles */, double* out, cudaStream_t st) {
  long long g = (n + 255) / 256;
  if (g > kSseCtas) g = kSseCtas;
  if (g < 1) g = 1;
  mean_sqerr_kernel<<<(unsigned)g, 256, 0, st>>>(xa, ld, k, n, truth, partial);
  sum_partials_kernel<<<1, 32, 0, st>>>(partial, (int)g, out);
  return cudaGetLastError();
}

cudaError_t transform_launch(const TransformArgs& a, int sms, cudaStream_t st) {
  const int k8 = (a.k + 7) / 8 <= 2 ? 2 : ((a.k + 7) / 8 <= 4 ? 4 : ((a.k + 7) / 8 <= 8 ? 8 : 16));
  const long long n_cols = a.n_loc + a.halo_l + a.halo_r;
  const long long n_tiles = (n_cols + TC - 1) / TC;
  const size_t smem = (size_t)8 * k8 * XLD * sizeof(double);
  long long g = n_tiles;                       // small problems: one 64-column chunk per CTA
  static int occ_cache[17] = {0};              // resident CTAs per SM of transform_kernel<K8>
  cudaError_t launch_err = cudaSuccess;
#define DAB_TR_CASE(K)                                                                              \
  {                                                                                                 \
    if (occ_cache[K] == 0) {                                                                        \
      cudaFuncSetAttribute(transform_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      int occ = 0;                                                                                  \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, transform_kernel<K>, 32 * K, smem);       \
      occ_cache[K] = occ < 1 ? 1 : occ;                                                             \
    }                                                                                               \
    const long long cap = (long long)sms * occ_cache[K];   /* every resident slot gets an equal share */ \
    if (g > cap) g = cap;                                                                           \
    if (g < 1) g = 1;                                                                               \
    launch_err = launch_chain(true, transform_kernel<K>, dim3((unsigned)g), dim3(32 * K), smem, st, a);    \
  }
  switch (k8) {
    case 2: DAB_TR_CASE(2) break;
    case 4: DAB_TR_CASE(4) break;
    case 8: DAB_TR_CASE(8) break;
    default: DAB_TR_CASE(16) break;
  }
#undef DAB_TR_CASE
  return launch_err != cudaSuccess ? launch_err : cudaGetLastError();
}

}  // namespace dab
