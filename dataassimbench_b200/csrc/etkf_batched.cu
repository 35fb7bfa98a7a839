// Batched small-problem ETKF cycling: one CTA runs one whole experiment (all cycles) on chip.
//
// BASELINE.json configs[1]: "4096 independent Lorenz96 N=40, k=20 ETKF experiments batched on 1
// B200 (inflation sweep)".  Every experiment is the reference's `ETKF.cycle()`
// (dabench/dacycler/_dacycler.py:186-290: analysis -> forecast, n_cycles times) with its own
// initial ensemble, observations and multiplicative inflation rho.  At these sizes a cycle is a
// few thousand flops, so the per-kernel path (5 launches per cycle) would be launch-bound; here
// the ensemble (with periodic halo copies), the k x k statistics, the Jacobi eigensolver and the
// RK4 window all live in one CTA's shared memory / registers and the experiments spread over the
// SMs with no communication at all ("replicas only").
//
// Same arithmetic as the large path: gather -> centred contraction -> transform matrix T (in-CTA
// Newton-Schulz, etkf_ns_device.cuh; Jacobi, etkf_eig_device.cuh, as its fallback) -> Xa = Xb T ->
// RK4 (fixed step) window.
#include <cmath>

#include "common.cuh"
#include "etkf_eig_device.cuh"
#include "etkf_ns_device.cuh"
#include "kernels.h"

namespace dab {

template <int PPT>    // points per thread: ceil(k*N / 256)
__global__ void __launch_bounds__(256)
etkf_batched_kernel(const BatchedArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int k = a.k, N = a.N, NP = N + 3;              // member row: [x_{N-2} x_{N-1} | x_0..x_{N-1} | x_0]
  const int kN = k * N, nyp = a.t_pad * a.n_obs;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  double* S0 = sm;                                      // k*NP  state / RK stage buffers (ping-pong)
  double* S1 = S0 + k * NP;
  double* Ys = S1 + k * NP;                             // k*nyp centred Yb^T
  double* dv = Ys + k * nyp;                            // nyp   innovations
  double* stats = dv + nyp;                             // k*k + k
  double* Tm = stats + k * k + k;                       // k*k
  double* ew = Tm + k * k;                              // eigensolver workspace

  const double rho = a.rho[b];
  const double r_inv = 1.0 / (a.sd * a.sd);
  const double inv_k = 1.0 / (double)k;
  const double dt = a.dt, F = a.F, h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0;
  const int lpp = eig_small_lpp(k);

  // this thread's grid points: p = tid + 256*q  ->  member j, column i, offset o into a state buffer
  int off[PPT], mem[PPT], col[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + 256 * q;
    const int j = p < kN ? p / N : 0;
    mem[q] = j; col[q] = p - j * N; off[q] = p < kN ? j * NP + 2 + (p - j * N) : -1;
  }
  auto put = [&](double* S, int q, double v) {          // store a point and its periodic halo copies
    const int o = off[q], i = col[q], base = mem[q] * NP;
    S[o] = v;
    if (i >= N - 2) S[base + i - (N - 2)] = v;
    if (i == 0) S[base + N + 2] = v;
  };

  const double* x0 = a.x0 + (long long)b * kN;
  double* cur = S0;
  double* nxt = S1;
#pragma unroll
  for (int q = 0; q < PPT; ++q)
    if (off[q] >= 0) put(cur, q, x0[tid + 256 * q]);
  __syncthreads();

  const double* vals_b = a.obs_val + (long long)b * a.n_obs_times * a.n_obs;
  const long long* idx_b = a.obs_idx + (long long)b * a.idx_batch_stride;

  for (int c = 0; c < a.n_cycles; ++c) {
    // ---- observations of this window -> centred Yb^T tile and innovations
    for (int r = tid; r < nyp; r += 256) {
      const int s = r / a.n_obs, o = r - s * a.n_obs;
      const long long ti = a.padded[(long long)c * a.t_pad + s];
      double y = 0.0; int loc = -1;
      if (ti > 0) {
        y = vals_b[(ti - 1) * a.n_obs + o];
        loc = (int)idx_b[(ti - 1) * a.n_obs + o];
        if (y != y) { loc = -1; y = 0.0; }              // NaN = padded slot of a non-stationary network
      }
      double sum = 0.0;
      for (int j = 0; j < k; ++j) {
        const double v = loc >= 0 ? cur[j * NP + 2 + loc] : 0.0;
        Ys[j * nyp + r] = v;
        sum += v;
      }
      const double ybar = sum * inv_k;
      for (int j = 0; j < k; ++j) Ys[j * nyp + r] = loc >= 0 ? Ys[j * nyp + r] - ybar : 0.0;
      dv[r] = loc >= 0 ? y - ybar : 0.0;
    }
    __syncthreads();
    // ---- C = Y'^T Rinv Y',  g = Y'^T Rinv d
    for (int e = tid; e < k * k + k; e += 256) {
      double s = 0.0;
      if (e < k * k) {
        const int p = e / k, q = e - p * k;
        for (int r = 0; r < nyp; ++r) s = fma(Ys[p * nyp + r], Ys[q * nyp + r], s);
      } else {
        const int p = e - k * k;
        for (int r = 0; r < nyp; ++r) s = fma(Ys[p * nyp + r], dv[r], s);
      }
      stats[e] = s * r_inv;
    }
    __syncthreads();
    // ---- k x k transform matrix T (shared memory): Newton-Schulz inverse square root on the tensor
    //      core; the Jacobi eigensolver for the empty-R branch, ill-conditioned or non-finite statistics
    if (a.empty_R || a.force_jacobi || !ns_small_body(stats, k, rho, Tm, nullptr, ew)) {
      if (lpp == 1) eig_small_body<1>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
      else if (lpp == 2) eig_small_body<2>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
      else eig_small_body<4>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
    }
    __syncthreads();
    // ---- Xa = Xb T for this thread's points
    double x[PPT], acc[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      double v = 0.0;
      if (off[q] >= 0) {
        const int j = mem[q], i = col[q];
        for (int m = 0; m < k; ++m) v = fma(Tm[m * k + j], cur[m * NP + 2 + i], v);
      }
      x[q] = v;
    }
    if (a.out_analysis != nullptr) {
      double* oa = a.out_analysis + ((long long)b * a.n_cycles + c) * kN;
#pragma unroll
      for (int q = 0; q < PPT; ++q)
        if (off[q] >= 0) oa[tid + 256 * q] = x[q];
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (off[q] >= 0) put(nxt, q, x[q]);
    __syncthreads();
    { double* t_ = cur; cur = nxt; nxt = t_; }
    if (a.out_mean != nullptr) {                          // ensemble-mean analysis (for RMSE scoring)
      double* om = a.out_mean + ((long long)b * a.n_cycles + c) * N;
      for (int i = tid; i < N; i += 256) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += cur[j * NP + 2 + i];
        om[i] = s * inv_k;
      }
    }
    // ---- RK4 window: stage inputs ping-pong through shared memory, x / acc stay in registers
    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const double ca = (st == 2) ? dt : h2;
        const double cw = (st == 0 || st == 3) ? h6 : h3;
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          if (off[q] >= 0) {
            const int o = off[q];
            const double kk = fma(cur[o + 1] - cur[o - 2], cur[o - 1], F - cur[o]);
            acc[q] = (st == 0) ? fma(cw, kk, x[q]) : fma(cw, kk, acc[q]);
            double v;
            if (st < 3) v = fma(ca, kk, x[q]);
            else { x[q] = acc[q]; v = acc[q]; }
            put(nxt, q, v);
          }
        }
        __syncthreads();
        { double* t_ = cur; cur = nxt; nxt = t_; }
      }
    }
  }
  if (a.out_final != nullptr) {
    double* of = a.out_final + (long long)b * kN;
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (off[q] >= 0) of[tid + 256 * q] = cur[off[q]];
  }
}

size_t batched_smem_bytes(int k, int N, int t_pad, int n_obs) {
  const int nyp = t_pad * n_obs;
  size_t d = 2 * (size_t)k * (N + 3) + (size_t)k * nyp + nyp + ((size_t)k * k + k) + (size_t)k * k +
             (size_t)(eig_small_smem_doubles(k) > ns_small_smem_doubles() ? eig_small_smem_doubles(k)
                                                                          : ns_small_smem_doubles());
  return d * sizeof(double) + 16;
}

cudaError_t batched_launch(const BatchedArgs& a, int batch, cudaStream_t st) {
  const size_t smem = batched_smem_bytes(a.k, a.N, a.t_pad, a.n_obs);
  const int ppt = (a.k * a.N + 255) / 256;
#define DAB_B_CASE(P)                                                                                 \
  {                                                                                                   \
    cudaFuncSetAttribute(etkf_batched_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    etkf_batched_kernel<P><<<batch, 256, smem, st>>>(a);                                              \
  }
  if (ppt <= 4) DAB_B_CASE(4)
  else if (ppt <= 8) DAB_B_CASE(8)
  else DAB_B_CASE(16)
#undef DAB_B_CASE
  return cudaGetLastError();
}

}  // namespace dab
