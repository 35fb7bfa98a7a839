// K2 + K3 -- observation gather and the ensemble-space contraction.
//
// Replaces (dabench/dacycler/_etkf.py, dabench/dacycler/_dacycler.py):
//   _calc_default_H (dense one-hot H)          _dacycler.py:58-66     -> the index list IS the operator
//   mask application on H                       _etkf.py:202-203       -> masked rows are dropped
//   _apply_obsop  Yb = H @ Xb                   _etkf.py:82-92, 131    -> gather (bit-exact copy)
//   yb_bar, Yb_pert                             _etkf.py:137-138
//   Yb_pert.T @ Rinv @ Yb_pert                  _etkf.py:144-146       -> C  (k x k)
//   Yb_pert.T @ Rinv @ (Y - yb_bar)             _etkf.py:154           -> g  (k)
// with R = sigma^2 I (`_calc_default_R`, _dacycler.py:68-72), so Rinv is the scalar 1/sigma^2
// applied once in the reduction.
//
// K3: a CTA walks blocks of 32 observation rows; the k x 32 tile Yb^T is gathered straight
// from the state (never materialised in HBM), centred in registers, parked in shared memory
// and contracted with DMMA.8x8x4 (FP64 tensor core): warp w owns the 8-row strip w of C.
// Every CTA writes its partial (C, g) to a workspace; `stats_reduce_kernel` sums the partials
// in a fixed order (bit-reproducible, no atomics) and applies 1/sigma^2.
#include "common.cuh"
#include "kernels.h"

namespace dab {

constexpr int RB = 32;           // observation rows per block
constexpr int YLD = RB + 4;      // smem row stride (doubles): (4*j + r) mod 16 distinct -> conflict-free frags

// One analysis's observation rows: up to `n_seg` segments of a CSR store.
//   seg[s] = {begin (into obs_loc/obs_val), count, first row number}
template <int K8>
__global__ void __launch_bounds__(32 * K8)
contract_kernel(const double* __restrict__ X, long long ld, int k,
                const int* __restrict__ obs_loc, const double* __restrict__ obs_val,
                const long long* __restrict__ seg, int n_seg, long long n_rows,
                double* __restrict__ partial) {
  constexpr int KP = 8 * K8;
  __shared__ double Ys[KP * YLD];     // centred Yb^T tile [member][row]
  __shared__ double psum[K8][RB];     // per-warp partial row sums
  __shared__ double dvec[RB];         // innovation y - ybar (0 for padding rows)
  __shared__ int s_loc[RB];
  __shared__ double s_val[RB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv_k = 1.0 / (double)k;

  double c0[K8], c1[K8];
#pragma unroll
  for (int t = 0; t < K8; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
  double gacc = 0.0;

  const long long n_blocks = (n_rows + RB - 1) / RB;
  for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    __syncthreads();                                   // previous block's tile fully consumed
    if (tid < RB) {
      const long long r = b * RB + tid;
      int loc = -1; double val = 0.0;
      if (r < n_rows) {
        int s = 0;
        while (s + 1 < n_seg && seg[3 * (s + 1) + 2] <= r) ++s;
        const long long p = seg[3 * s] + (r - seg[3 * s + 2]);
        loc = obs_loc[p]; val = obs_val[p];
      }
      s_loc[tid] = loc; s_val[tid] = val;
    }
    __syncthreads();
    // gather: this thread owns row `lane`, members warp, warp+K8, ... (8 of them)
    const int loc = s_loc[lane];
    double v[8];
    double ps = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      v[i] = (loc >= 0 && j < k) ? X[(long long)j * ld + loc] : 0.0;
      ps += v[i];
    }
    psum[warp][lane] = ps;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < K8; ++w) tot += psum[w][lane];
    const double ybar = tot * inv_k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      Ys[j * YLD + lane] = (loc >= 0 && j < k) ? v[i] - ybar : 0.0;
    }
    if (warp == 0) dvec[lane] = (loc >= 0) ? s_val[lane] - ybar : 0.0;
    __syncthreads();
    // C strip `warp` += Y'^T Y' over the 32 rows, 4 rows per DMMA
    const int fm = lane >> 2, fr = lane & 3;
#pragma unroll
    for (int r0 = 0; r0 < RB; r0 += 4) {
      const double a = Ys[(8 * warp + fm) * YLD + r0 + fr];
#pragma unroll
      for (int t = 0; t < K8; ++t) {
        const double bb = Ys[(8 * t + fm) * YLD + r0 + fr];
        dmma884(c0[t], c1[t], a, bb);
      }
    }
    if (tid < KP) {
      double gs = 0.0;
#pragma unroll 8
      for (int r = 0; r < RB; ++r) gs = fma(Ys[tid * YLD + r], dvec[r], gs);
      gacc += gs;
    }
  }
  // partial layout: [KP][KP] row-major then [KP]
  double* out = partial + (long long)blockIdx.x * (KP * KP + KP);
  const int row = 8 * warp + (lane >> 2);
#pragma unroll
  for (int t = 0; t < K8; ++t) {
    const int col = 8 * t + 2 * (lane & 3);
    out[row * KP + col] = c0[t];
    out[row * KP + col + 1] = c1[t];
  }
  if (tid < KP) out[KP * KP + tid] = gacc;
}

// stats[0 .. k*k) = r_inv * sum_p C_p (row-major k x k), stats[k*k .. k*k+k) = r_inv * sum_p g_p.
// 8 slices per element, each summing its partials in order; slices combined in order.
__global__ void __launch_bounds__(256)
stats_reduce_kernel(const double* __restrict__ partial, int n_part, int kp, int k, double r_inv,
                    double* __restrict__ stats) {
  __shared__ double sl[8][32];
  const int el = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  const int n_el = k * k + k;
  double s = 0.0;
  if (el < n_el) {
    const int src = el < k * k ? (el / k) * kp + (el % k) : kp * kp + (el - k * k);
    const long long stride = (long long)kp * kp + kp;
    for (int p = slice; p < n_part; p += 8) s += partial[(long long)p * stride + src];
  }
  sl[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && el < n_el) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sl[i][threadIdx.x];
    stats[el] = t * r_inv;
  }
}

int contract_k8(int k) { return k <= 16 ? 2 : (k <= 32 ? 4 : (k <= 64 ? 8 : 16)); }

size_t contract_partial_doubles(int k, int n_cta) {
  const int kp = 8 * contract_k8(k);
  return (size_t)n_cta * ((size_t)kp * kp + kp);
}

int contract_grid(long long n_rows, int sms) {
  const long long nb = (n_rows + RB - 1) / RB;
  long long g = nb < 1 ? 1 : nb;
  if (g > sms) g = sms;
  return (int)g;
}

cudaError_t contract_launch(const double* X, long long ld, int k, const int* obs_loc,
                            const double* obs_val, const long long* seg, int n_seg,
                            long long n_rows, double* partial, int grid, cudaStream_t st) {
  const int k8 = contract_k8(k);
  switch (k8) {
    case 2: contract_kernel<2><<<grid, 64, 0, st>>>(X, ld, k, obs_loc, obs_val, seg, n_seg, n_rows, partial); break;
    case 4: contract_kernel<4><<<grid, 128, 0, st>>>(X, ld, k, obs_loc, obs_val, seg, n_seg, n_rows, partial); break;
    case 8: contract_kernel<8><<<grid, 256, 0, st>>>(X, ld, k, obs_loc, obs_val, seg, n_seg, n_rows, partial); break;
    default: contract_kernel<16><<<grid, 512, 0, st>>>(X, ld, k, obs_loc, obs_val, seg, n_seg, n_rows, partial); break;
  }
  return cudaGetLastError();
}

cudaError_t stats_reduce_launch(const double* partial, int n_part, int k, double r_inv, double* stats,
                                cudaStream_t st) {
  const int n_el = k * k + k;
  stats_reduce_kernel<<<(n_el + 31) / 32, 256, 0, st>>>(partial, n_part, 8 * contract_k8(k), k, r_inv, stats);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: Yb[j][r] = X[j][idx[r]] (0 where mask[r] == 0).  A pure copy, hence bit-exact.
__global__ void __launch_bounds__(256)
gather_kernel(const double* __restrict__ X, long long ld, int k, const long long* __restrict__ idx,
              const unsigned char* __restrict__ mask, long long n_y, double* __restrict__ Yb) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n_y) return;
  const long long loc = idx[r];
  const bool ok = mask == nullptr || mask[r] != 0;
  for (int j = blockIdx.y; j < k; j += gridDim.y)
    Yb[(long long)j * n_y + r] = ok ? X[(long long)j * ld + loc] : 0.0;
}

cudaError_t gather_launch(const double* X, long long ld, int k, const long long* idx,
                          const unsigned char* mask, long long n_y, double* Yb, cudaStream_t st) {
  if (n_y <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n_y + 255) / 256), (unsigned)(k < 64 ? k : 64));
  gather_kernel<<<grid, 256, 0, st>>>(X, ld, k, idx, mask, n_y, Yb);
  return cudaGetLastError();
}

// Build a compact (loc, val) list on device from reference-style arrays (used by the
// single-analysis entry point): rows with mask == 0 are kept as loc = -1 (contribute nothing).
__global__ void pack_obs_kernel(const long long* __restrict__ idx, const unsigned char* __restrict__ mask,
                                const double* __restrict__ y, long long n_y,
                                int* __restrict__ loc, double* __restrict__ val) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n_y) return;
  const bool ok = mask == nullptr || mask[r] != 0;
  loc[r] = ok ? (int)idx[r] : -1;
  val[r] = ok ? y[r] : 0.0;
}

cudaError_t pack_obs_launch(const long long* idx, const unsigned char* mask, const double* y,
                            long long n_y, int* loc, double* val, cudaStream_t st) {
  if (n_y <= 0) return cudaSuccess;
  pack_obs_kernel<<<(unsigned)((n_y + 255) / 256), 256, 0, st>>>(idx, mask, y, n_y, loc, val);
  return cudaGetLastError();
}

}  // namespace dab
