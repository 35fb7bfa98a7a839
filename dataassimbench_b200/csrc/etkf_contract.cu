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
// K3: a CTA walks its rows in blocks of 32; the k x 32 tile Yb^T is gathered straight from the
// state (never materialised in HBM), centred in registers, parked in shared memory and
// contracted with DMMA.8x8x4 (FP64 tensor core): warp w owns the 8-row strip w of C.
// Every CTA writes its partial (C, g) to a workspace; `stats_reduce_kernel` sums the partials
// in a fixed order (bit-reproducible, no atomics) and applies 1/sigma^2.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dab {

constexpr int RB = 32;           // observation rows per block
constexpr int YLD = RB + 4;      // smem row stride (doubles): (4*j + r) mod 16 distinct -> conflict-free frags

// One analysis's observation rows: up to `n_seg` segments of a CSR store.
//   seg[s] = {begin (into obs_loc/obs_val), count, first row number}
// Every CTA owns one contiguous range of rows (rows are ordered by grid location, so a CTA's
// gathers stay in one neighbourhood of the state) and walks it in blocks of RB rows, software-
// pipelined: the gather of block b+1 (L2/HBM latency) is in flight while the tensor core
// contracts block b, and the (location, value) pairs of block b+2 are fetched one block earlier
// still.  A short last block runs only the k-steps it has rows for.
// `obs_sw` (nullable): 1 / sigma of every row for a diagonal R with unequal entries (0 where sigma = 0: pinv);
// the centred row and its innovation are scaled by it, so C and g come out weighted by 1 / sigma^2.
__device__ __forceinline__ void fetch_row(const int* __restrict__ obs_loc, const double* __restrict__ obs_val,
                                          const double* __restrict__ obs_sw,
                                          const long long* __restrict__ seg, int n_seg, long long r,
                                          long long r_end, int& loc, double& val, double& sw) {
  loc = -1; val = 0.0; sw = 1.0;
  if (r < r_end) {
    int s = 0;
    while (s + 1 < n_seg && seg[3 * (s + 1) + 2] <= r) ++s;
    const long long p = seg[3 * s] + (r - seg[3 * s + 2]);
    loc = obs_loc[p]; val = obs_val[p];
    if (obs_sw != nullptr) sw = obs_sw[p];
  }
}

template <int K8>
__global__ void __launch_bounds__(32 * K8)
contract_kernel(const double* __restrict__ X, long long ld, int k,
                const int* __restrict__ obs_loc, const double* __restrict__ obs_val,
                const double* __restrict__ obs_sw,
                const long long* __restrict__ seg, int n_seg, long long n_rows,
                double* __restrict__ partial) {
  constexpr int KP = 8 * K8;
  __shared__ double Ys[KP * YLD];     // centred Yb^T tile [member][row]
  __shared__ double psum[K8][RB];     // per-warp partial row sums
  __shared__ double dvec[RB];         // innovation y - ybar (0 for padding rows)
  __shared__ int s_loc[2][RB];
  __shared__ double s_val[2][RB], s_sw[2][RB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv_k = 1.0 / (double)k;

  double c0[K8], c1[K8];
#pragma unroll
  for (int t = 0; t < K8; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
  double gacc = 0.0;

  // contiguous, balanced row range of this CTA
  const long long per = n_rows / gridDim.x, extra = n_rows % gridDim.x;
  const long long r_begin = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const long long r_end = r_begin + per + (blockIdx.x < extra ? 1 : 0);
  const int n_blocks = (int)((r_end - r_begin + RB - 1) / RB);

  // pipeline prologue: rows of block 0 -> shared, rows of block 1 -> registers, gather of block 0
  int nloc = -1; double nval = 0.0, nsw = 1.0;   // (tid < RB) the row `tid` of the block after next
  if (tid < RB) {
    fetch_row(obs_loc, obs_val, obs_sw, seg, n_seg, r_begin + tid, r_end, nloc, nval, nsw);
    s_loc[0][tid] = nloc; s_val[0][tid] = nval; s_sw[0][tid] = nsw;
    fetch_row(obs_loc, obs_val, obs_sw, seg, n_seg, r_begin + RB + tid, r_end, nloc, nval, nsw);
  }
  __syncthreads();
  // everything above read the observation tables only; the state is the previous kernel's output
  pdl_wait();
  pdl_launch_dependents();
  double v[8];
  int loc = s_loc[0][lane];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = warp + i * K8;
    v[i] = (loc >= 0 && j < k) ? X[(long long)j * ld + loc] : 0.0;
  }

  for (int b = 0; b < n_blocks; ++b) {
    const int cur = b & 1;
    // ---- centre block b (its gather has landed in v) and park it in shared memory
    double ps = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) ps += v[i];
    psum[warp][lane] = ps;
    if (tid < RB) { s_loc[cur ^ 1][tid] = nloc; s_val[cur ^ 1][tid] = nval; s_sw[cur ^ 1][tid] = nsw; }   // rows of block b+1
    __syncthreads();                                   // also: every warp is done with tile b-1
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < K8; ++w) tot += psum[w][lane];
    const double ybar = tot * inv_k;
    const double sw = s_sw[cur][lane];                 // 1 unless R has unequal diagonal entries (x * 1.0 is exact)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      Ys[j * YLD + lane] = (loc >= 0 && j < k) ? (v[i] - ybar) * sw : 0.0;
    }
    if (warp == 0) dvec[lane] = (loc >= 0) ? (s_val[cur][lane] - ybar) * sw : 0.0;
    __syncthreads();
    // ---- put the next gather (block b+1) and the rows of block b+2 in flight
    loc = s_loc[cur ^ 1][lane];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      v[i] = (loc >= 0 && j < k) ? X[(long long)j * ld + loc] : 0.0;
    }
    if (tid < RB) fetch_row(obs_loc, obs_val, obs_sw, seg, n_seg, r_begin + (long long)(b + 2) * RB + tid, r_end, nloc, nval, nsw);
    // ---- C strip `warp` += Y'^T Y' over the rows of block b, 4 rows per DMMA
    const long long left = r_end - (r_begin + (long long)b * RB);
    const int rows_here = left < RB ? (int)left : RB;
    const int fm = lane >> 2, fr = lane & 3;
#pragma unroll 2
    for (int r0 = 0; r0 < rows_here; r0 += 4) {
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
    *reinterpret_cast<double2*>(out + row * KP + col) = make_double2(c0[t], c1[t]);
  }
  if (tid < KP) out[KP * KP + tid] = gacc;
}

// stats[0 .. k*k) = r_inv * sum_p C_p (row-major k x k), stats[k*k .. k*k+k) = r_inv * sum_p g_p.
// 16 slices per element, each summing its partials in a fixed order (four interleaved chains so
// the loads overlap); slices combined in order: bit-reproducible, no atomics.
//
// Grid-sharded runs (push.world > 1): the k*k+k statistics of all ranks have to be summed.  Instead
// of an NCCL all-reduce behind this kernel, the reduction is fused with an all-gather over peer
// memory: every rank stores its reduced values straight into slot `rank` of EVERY rank's exchange
// buffer (P2P stores over NVLink, 256 contiguous bytes per CTA and peer), and the last CTA to
// finish raises this rank's sequence flag on every rank.  `stats_gather_kernel` then waits for the
// flags of all ranks and adds the slots in rank order -- the same bits on every rank.
__global__ void __launch_bounds__(512)
stats_reduce_kernel(const double* __restrict__ partial, int n_part, int kp, int k, double r_inv,
                    double* __restrict__ stats, const StatsPush push) {
  constexpr int NS = 16;
  __shared__ double sl[NS][32];
  const int el = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  const int n_el = k * k + k;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  pdl_wait();
  pdl_launch_dependents();
  if (el < n_el) {
    const int src = el < k * k ? (el / k) * kp + (el % k) : kp * kp + (el - k * k);
    const long long stride = (long long)kp * kp + kp;
    const double* pp = partial + src;
    int p = slice;
    for (; p + 3 * NS < n_part; p += 4 * NS) {
      s0 += pp[(long long)p * stride];
      s1 += pp[(long long)(p + NS) * stride];
      s2 += pp[(long long)(p + 2 * NS) * stride];
      s3 += pp[(long long)(p + 3 * NS) * stride];
    }
    if (p < n_part) s0 += pp[(long long)p * stride];
    if (p + NS < n_part) s1 += pp[(long long)(p + NS) * stride];
    if (p + 2 * NS < n_part) s2 += pp[(long long)(p + 2 * NS) * stride];
  }
  sl[slice][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice == 0 && el < n_el) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) t += sl[i][threadIdx.x];
    t *= r_inv;
    if (push.world <= 1) {
      stats[el] = t;
    } else {
      for (int g = 0; g < push.world; ++g) push.slot[g][el] = t;
      __threadfence_system();
    }
  }
  if (push.world > 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(push.done_ctr, 1u);
      if (prev == gridDim.x - 1) {               // every CTA of this rank has pushed (and fenced)
        *push.done_ctr = 0;
        __threadfence_system();
        for (int g = 0; g < push.world; ++g) *reinterpret_cast<volatile unsigned long long*>(push.flag[g]) = push.seq;
      }
    }
  }
}

// Waits until every rank's flag has reached `seq` (P2P stores, see above), then
// stats = sum over ranks of their slots, in rank order.  A peer that never arrives is a lost
// process: after `timeout_ns` (10 s by default, DAB_XCHG_TIMEOUT_S; 0 = wait for ever, e.g. under a
// profiler's kernel replay or a debugger) the kernel traps instead of hanging the GPU.
__global__ void __launch_bounds__(256)
stats_gather_kernel(const unsigned long long* flags, unsigned long long seq, const double* slots,
                    int n_el, int world, double* __restrict__ stats, unsigned long long timeout_ns) {
  if (threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int g = 0; g < world; ++g) {
      const volatile unsigned long long* f = flags + g;
      while (*f < seq) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (timeout_ns != 0ull && t1 - t0 > timeout_ns) __trap();
        __nanosleep(100);
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  // the flags (this rank's own included) are what this kernel really depends on, so the polling
  // above may overlap the tail of the reduction kernel; the wait keeps the stream-order chain
  pdl_wait();
  pdl_launch_dependents();
  const int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el < n_el) {
    double s = 0.0;
    for (int g = 0; g < world; ++g) s += __ldcg(slots + (long long)g * n_el + el);
    stats[el] = s;
  }
}

int contract_k8(int k) { return k <= 16 ? 2 : (k <= 32 ? 4 : (k <= 64 ? 8 : 16)); }

size_t contract_partial_doubles(int k, int n_cta) {
  const int kp = 8 * contract_k8(k);
  return (size_t)n_cta * ((size_t)kp * kp + kp);
}

int contract_grid(long long n_rows, int sms) {
  static int mult = 0;                 // experiment knob: CTAs per SM the row ranges are cut for (default 1)
  if (mult == 0) { const char* ev = getenv("DAB_K3_CTAS_PER_SM"); mult = ev ? atoi(ev) : 1; if (mult < 1 || mult > 4) mult = 1; }
  const long long nb = (n_rows + RB - 1) / RB;
  long long g = nb < 1 ? 1 : nb;
  if (g > (long long)sms * mult) g = (long long)sms * mult;
  return (int)g;
}

cudaError_t contract_launch(const double* X, long long ld, int k, const int* obs_loc,
                            const double* obs_val, const double* obs_sw, const long long* seg, int n_seg,
                            long long n_rows, double* partial, int grid, bool chain, cudaStream_t st) {
  const int k8 = contract_k8(k);
  switch (k8) {
    case 2: return launch_chain(chain, contract_kernel<2>, dim3(grid), dim3(64), 0, st, X, ld, k, obs_loc, obs_val, obs_sw, seg, n_seg, n_rows, partial);
    case 4: return launch_chain(chain, contract_kernel<4>, dim3(grid), dim3(128), 0, st, X, ld, k, obs_loc, obs_val, obs_sw, seg, n_seg, n_rows, partial);
    case 8: return launch_chain(chain, contract_kernel<8>, dim3(grid), dim3(256), 0, st, X, ld, k, obs_loc, obs_val, obs_sw, seg, n_seg, n_rows, partial);
    default: return launch_chain(chain, contract_kernel<16>, dim3(grid), dim3(512), 0, st, X, ld, k, obs_loc, obs_val, obs_sw, seg, n_seg, n_rows, partial);
  }
  return cudaGetLastError();
}

cudaError_t stats_reduce_launch(const double* partial, int n_part, int k, double r_inv, double* stats,
                                const StatsPush* push, cudaStream_t st) {
  const int n_el = k * k + k;
  StatsPush none{};
  none.world = 1;
  return launch_chain(true, stats_reduce_kernel, dim3((n_el + 31) / 32), dim3(512), 0, st, partial, n_part,
                      8 * contract_k8(k), k, r_inv, stats, push ? *push : none);
}

cudaError_t stats_gather_launch(const unsigned long long* flags, unsigned long long seq, const double* slots,
                                int k, int world, double* stats, unsigned long long timeout_ns, cudaStream_t st) {
  const int n_el = k * k + k;
  return launch_chain(true, stats_gather_kernel, dim3((n_el + 255) / 256), dim3(256), 0, st, flags, seq, slots,
                      n_el, world, stats, timeout_ns);
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

// ------------------------------------------------------------------------------------------
// Observer.observe sampling (the step before the path; dabench/observer/_observer.py:290-316,
// 346-363): values[t][j] = state[time_pos[t]][loc[t][j]] + errors[t][j] for j < count[t]; the
// ragged tail of a row is NaN (the reference pads with NaN, and NaN + error stays NaN).  The
// index draws stay on the host (they must consume NumPy's PCG64 stream in the reference's order);
// this is the part that touches the trajectory, so a device-resident nature run never has to be
// copied out: 8 B read (one 32 B sector) + 8 B index + 8 B error + 8 B write per observation.
// An index outside [0, N) yields NaN and raises *bad (checked by the host wrapper).
__global__ void __launch_bounds__(256)
observe_kernel(const double* __restrict__ state, long long N, long long n_state_times,
               const long long* __restrict__ time_pos, long long n_t, const long long* __restrict__ loc, int stationary,
               const long long* __restrict__ count, long long n_obs, const double* __restrict__ errors,
               double* __restrict__ values, int* __restrict__ bad) {
  const long long total = n_t * n_obs;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / n_obs, j = i - t * n_obs;
    double v = nan;
    if (count == nullptr || j < count[t]) {
      const long long l = loc[stationary ? j : i], tp = time_pos[t];
      if (l >= 0 && l < N && tp >= 0 && tp < n_state_times) {
        v = state[tp * N + l];
        if (errors != nullptr) v += errors[i];
      } else {
        *bad = 1;
      }
    }
    values[i] = v;
  }
}

cudaError_t observe_launch(const double* state, long long N, long long n_state_times, const long long* time_pos, long long n_t,
                           const long long* loc, int stationary, const long long* count, long long n_obs,
                           const double* errors, double* values, int* bad, int sms, cudaStream_t st) {
  const long long total = n_t * n_obs;
  if (total <= 0) return cudaSuccess;
  long long g = (total + 255) / 256;
  const long long cap = (long long)sms * 8;          // 8 resident 256-thread CTAs per SM, grid-stride beyond
  if (g > cap) g = cap;
  observe_kernel<<<(unsigned)g, 256, 0, st>>>(state, N, n_state_times, time_pos, n_t, loc, stationary, count, n_obs, errors,
                                              values, bad);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// 3D-Var analysis with the default operators.  Var3D._cycle_obsop
// (dabench/dacycler/_var3d.py:84-110) solves (I + B H^T R^-1 H) xa = xb + B H^T R^-1 y by conjugate
// gradients; with the default H (index gather, masked rows dropped), R = sd^2 I and B = I that system
// is diagonal: xa_i = (xb_i + sum_r w y_r) / (1 + sum_r w) over the valid rows r observing point i,
// w = 1/sd^2.  Three small HBM-bound kernels: copy + ones, scatter-add of the rows (FP64 atomics:
// a grid point observed once or twice gets the same bits in any order), divide.
__global__ void __launch_bounds__(256)
var3d_init_kernel(const double* __restrict__ xb, long long n, double* __restrict__ num, double* __restrict__ den) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    num[i] = xb[i];
    den[i] = 1.0;
  }
}

__global__ void __launch_bounds__(256)
var3d_scatter_kernel(const long long* __restrict__ loc, const double* __restrict__ val, long long n_rows,
                     long long n, double w, double* __restrict__ num, double* __restrict__ den,
                     int* __restrict__ bad) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long l = loc[r];
    if (l < 0 || l >= n) { *bad = 1; continue; }
    atomicAdd(num + l, w * val[r]);
    atomicAdd(den + l, w);
  }
}

__global__ void __launch_bounds__(256)
var3d_divide_kernel(double* __restrict__ num, const double* __restrict__ den, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    num[i] = num[i] / den[i];
}

cudaError_t var3d_analysis_launch(const double* xb, double* xa, double* den, long long n, const long long* loc,
                                  const double* val, long long n_rows, double w, int* bad, int sms,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long cap = (long long)sms * 8;
  long long g = (n + 255) / 256;
  if (g > cap) g = cap;
  var3d_init_kernel<<<(unsigned)g, 256, 0, st>>>(xb, n, xa, den);
  if (n_rows > 0) {
    long long gr = (n_rows + 255) / 256;
    if (gr > cap) gr = cap;
    var3d_scatter_kernel<<<(unsigned)gr, 256, 0, st>>>(loc, val, n_rows, n, w, xa, den, bad);
  }
  var3d_divide_kernel<<<(unsigned)g, 256, 0, st>>>(xa, den, n);
  return cudaGetLastError();
}

// Build a compact (loc, val) list on device from reference-style arrays (used by the
// single-analysis entry point): rows with mask == 0 are kept as loc = -1 (contribute nothing).
// An index of a valid row outside [0, N) is reported through *bad (and contributes nothing) instead of
// being dereferenced by the contraction.
__global__ void pack_obs_kernel(const long long* __restrict__ idx, const unsigned char* __restrict__ mask,
                                const double* __restrict__ y, const double* __restrict__ sd, long long n_y, long long N,
                                int* __restrict__ loc, double* __restrict__ val, double* __restrict__ sw,
                                int* __restrict__ bad) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n_y) return;
  bool ok = mask == nullptr || mask[r] != 0;
  const long long l = idx[r];
  if (ok && (l < 0 || l >= N)) { ok = false; *bad = 1; }
  loc[r] = ok ? (int)l : -1;
  val[r] = ok ? y[r] : 0.0;
  if (sw != nullptr) { const double s = sd[r]; sw[r] = s == 0.0 ? 0.0 : 1.0 / s; }   // pinv of a zero variance
}

// Observation values that are already in HBM -> the cycler's location-ordered table: row i of time t is
// values[t][src[src_base[t] + i]]; with `mask_nan` (moving observers: NaN = padded slot, the location mask of
// dabench/dacycler/_dacycler.py:262-268) such a row gets location -1 and contributes nothing.
__global__ void __launch_bounds__(256)
pack_rows_kernel(const double* __restrict__ values, long long n_obs, const int* __restrict__ src,
                 const long long* __restrict__ off, const long long* __restrict__ src_base, int mask_nan,
                 int* __restrict__ loc, double* __restrict__ val) {
  const long long t = blockIdx.y;
  const long long n = off[t + 1] - off[t];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = values[t * n_obs + src[src_base[t] + i]];
    const bool drop = mask_nan && v != v;
    val[off[t] + i] = drop ? 0.0 : v;
    if (drop) loc[off[t] + i] = -1;
  }
}

cudaError_t pack_rows_launch(const double* values, long long n_obs, long long n_times, const int* src,
                             const long long* off, const long long* src_base, int mask_nan, int* loc, double* val,
                             cudaStream_t st) {
  if (n_obs <= 0 || n_times <= 0) return cudaSuccess;
  long long gx = (n_obs + 255) / 256;
  if (gx > 64) gx = 64;
  for (long long t0 = 0; t0 < n_times; t0 += 65535) {            // gridDim.y limit
    const long long nt = n_times - t0 < 65535 ? n_times - t0 : 65535;
    pack_rows_kernel<<<dim3((unsigned)gx, (unsigned)nt), 256, 0, st>>>(values + t0 * n_obs, n_obs, src, off + t0,
                                                                      src_base + t0, mask_nan, loc, val);
  }
  return cudaGetLastError();
}

// the first two stages alone: num = xb + sum_r w y_r, den = 1 + sum_r w (var3d_cg.cu takes its diagonal from them)
cudaError_t var3d_scatter_only_launch(const double* xb, double* num, double* den, long long n, const long long* loc,
                                      const double* val, long long n_rows, double w, int* bad, int sms, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long cap = (long long)sms * 8;
  long long g = (n + 255) / 256;
  if (g > cap) g = cap;
  var3d_init_kernel<<<(unsigned)g, 256, 0, st>>>(xb, n, num, den);
  if (n_rows > 0) {
    long long gr = (n_rows + 255) / 256;
    if (gr > cap) gr = cap;
    var3d_scatter_kernel<<<(unsigned)gr, 256, 0, st>>>(loc, val, n_rows, n, w, num, den, bad);
  }
  return cudaGetLastError();
}

cudaError_t pack_obs_launch(const long long* idx, const unsigned char* mask, const double* y, const double* sd,
                            long long n_y, long long N, int* loc, double* val, double* sw, int* bad, cudaStream_t st) {
  if (n_y <= 0) return cudaSuccess;
  pack_obs_kernel<<<(unsigned)((n_y + 255) / 256), 256, 0, st>>>(idx, mask, y, sd, n_y, N, loc, val, sd ? sw : nullptr, bad);
  return cudaGetLastError();
}

}  // namespace dab
