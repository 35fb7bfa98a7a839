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

// -DDAB_BATCH_TIMING (debug builds): thread 0 of CTA 0 accumulates clock64() per section and prints them
#ifdef DAB_BATCH_TIMING
#define BT_TICK(i) do { const long long now_ = clock64(); bt_[i] += now_ - bt_tick; bt_tick = now_; } while (0)
#else
#define BT_TICK(i) do { } while (0)
#endif

__host__ __device__ inline int batched_state_doubles(int k, int N, int ppt, bool blocked) {
  return blocked ? ppt * 256 : k * (N + 3);
}

// PPT points per thread.  BLOCKED: a thread owns PPT CONSECUTIVE points of one member
// (ceil(N/PPT) threads per member, k * ceil(N/PPT) <= 256) and keeps them in registers through the
// RK4 window, so a stage reads only 3 neighbour values from shared memory.  Otherwise (shapes that do
// not fit that mapping) points are dealt round-robin, p = tid + 256 q, and a stage reads all 4 stencil
// values of every point from shared memory.
template <int PPT, bool BLOCKED>
__global__ void __launch_bounds__(256)
etkf_batched_kernel(const BatchedArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int k = a.k, N = a.N, NP = N + 3;
  const int kN = k * N, nyp = a.t_pad * a.n_obs, nys = nyp | 1;   // odd row stride of Ys: conflict-free columns
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int tpm = (N + PPT - 1) / PPT;                  // BLOCKED: threads per member
  // State layout in shared memory.  BLOCKED: thread-major -- point q of thread t at q*256 + t, so
  // that the stores of a stage and the three neighbour loads are conflict-free (member-major rows
  // put consecutive threads PPT doubles apart: 4-way conflicts, 2.1e9 excess wavefronts in the
  // first ncu capture).  Otherwise member-major rows of N + 3.
  auto sidx = [&](int j, int i) -> int {
    return BLOCKED ? (i % PPT) * 256 + j * tpm + i / PPT : j * NP + 2 + i;
  };
  const int state_doubles = batched_state_doubles(k, N, PPT, BLOCKED);
  double* S0 = sm;                                      // state / RK stage buffers (ping-pong)
  double* S1 = S0 + state_doubles;
  double* Ys = S1 + state_doubles;                      // k*nyp centred Yb^T
  double* dv = Ys + k * nys;                            // nyp   innovations
  double* stats = dv + nyp;                             // k*k + k
  double* Tm = stats + k * k + k;                       // k*k
  double* ew = Tm + k * k;                              // eigensolver workspace

  const double rho = a.rho[b];
  const double r_inv = a.r_inv;
  const double inv_k = 1.0 / (double)k;
  const double dt = a.dt, F = a.F, h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0;
  const int lpp = eig_small_lpp(k);

  // this thread's grid points: p = tid + 256*q  ->  member j, column i, offset o into a state buffer,
  // and the offsets of its three stencil neighbours on the periodic ring (computed once: the RK4
  // stage below is then 4 loads, the 17-instruction arithmetic of l96_forecast.cu and 1 store)
  int off[PPT], mem[PPT], col[PPT], gix[PPT], o_m2[PPT], o_m1[PPT], o_p1[PPT];
  int npts = 0;                                         // BLOCKED: valid points of this thread
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    int j, i;
    bool valid;
    if (BLOCKED) {
      j = tid / tpm; i = (tid - j * tpm) * PPT + q;
      valid = j < k && i < N;
      if (valid) npts = q + 1;
    } else {
      const int p = tid + 256 * q;
      valid = p < kN;
      j = valid ? p / N : 0; i = p - j * N;
    }
    if (!valid) { j = 0; i = 0; }
    mem[q] = j; col[q] = i; gix[q] = j * N + i; off[q] = valid ? sidx(j, i) : -1;
    o_m2[q] = sidx(j, i >= 2 ? i - 2 : i - 2 + N);
    o_m1[q] = sidx(j, i >= 1 ? i - 1 : N - 1);
    o_p1[q] = sidx(j, i + 1 < N ? i + 1 : 0);
  }
  auto put = [&](double* S, int q, double v) { S[off[q]] = v; };

  const double* x0 = a.x0 + (long long)b * kN;
  double* cur = S0;
  double* nxt = S1;
#pragma unroll
  for (int q = 0; q < PPT; ++q)
    if (off[q] >= 0) put(cur, q, x0[gix[q]]);
  __syncthreads();

  const double* vals_b = a.obs_val + (long long)b * a.val_batch_stride;
  const long long* idx_b = a.obs_idx + (long long)b * a.idx_batch_stride;

#ifdef DAB_BATCH_TIMING
  long long bt_[6] = {0, 0, 0, 0, 0, 0}, bt_tick = clock64();
#endif
  for (int c = 0; c < a.n_cycles; ++c) {
    BT_TICK(5);
    // ---- observations of this window -> centred Yb^T tile and innovations
    for (int r = tid; r < nyp; r += 256) {
      const int s = r / a.n_obs, o = r - s * a.n_obs;
      const long long ti = a.padded[(long long)c * a.t_pad + s];
      double y = 0.0; int loc = -1;
      if (ti > 0) {
        y = vals_b[(ti - 1) * a.n_obs + o];
        loc = (int)idx_b[(ti - 1) * a.n_obs + o];
        // NaN = padded slot of a non-stationary network (_dacycler.py:266-268); for a stationary one the
        // reference's location mask is all ones and a NaN value is data (it propagates), as in engine.cu
        if (!a.stationary && y != y) { loc = -1; y = 0.0; }
      }
      double sum = 0.0;
      for (int j = 0; j < k; ++j) {
        const double v = loc >= 0 ? cur[sidx(j, loc)] : 0.0;
        Ys[j * nys + r] = v;
        sum += v;
      }
      const double ybar = sum * inv_k;
      for (int j = 0; j < k; ++j) Ys[j * nys + r] = loc >= 0 ? Ys[j * nys + r] - ybar : 0.0;
      dv[r] = loc >= 0 ? y - ybar : 0.0;
    }
    __syncthreads();
    BT_TICK(0);
    // ---- C = Y'^T Rinv Y',  g = Y'^T Rinv d
    for (int e = tid; e < k * k + k; e += 256) {
      double s = 0.0;
      if (e < k * k) {
        const int p = e / k, q = e - p * k;
        for (int r = 0; r < nyp; ++r) s = fma(Ys[p * nys + r], Ys[q * nys + r], s);
      } else {
        const int p = e - k * k;
        for (int r = 0; r < nyp; ++r) s = fma(Ys[p * nys + r], dv[r], s);
      }
      stats[e] = s * r_inv;
    }
    __syncthreads();
    BT_TICK(1);
    // ---- k x k transform matrix T (shared memory): Newton-Schulz inverse square root on the tensor
    //      core; the Jacobi eigensolver for the empty-R branch, ill-conditioned or non-finite statistics
    if (a.empty_R || a.force_jacobi || !ns_small_body(stats, k, rho, Tm, nullptr, ew)) {
      if (lpp == 1) eig_small_body<1>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
      else if (lpp == 2) eig_small_body<2>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
      else eig_small_body<4>(stats, k, rho, a.empty_R, Tm, nullptr, nullptr, ew);
    }
    __syncthreads();
    BT_TICK(2);
    // ---- Xa = Xb T for this thread's points
    double x[PPT], acc[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      double v = 0.0;
      if (off[q] >= 0) {
        const int j = mem[q], i = col[q];
        for (int m = 0; m < k; ++m) v = fma(Tm[m * k + j], cur[sidx(m, i)], v);
      }
      x[q] = v;
    }
    if (a.out_analysis != nullptr) {
      double* oa = a.out_analysis + ((long long)b * a.n_cycles + c) * kN;
#pragma unroll
      for (int q = 0; q < PPT; ++q)
        if (off[q] >= 0) oa[gix[q]] = x[q];
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
        for (int j = 0; j < k; ++j) s += cur[sidx(j, i)];
        om[i] = s * inv_k;
      }
    }
    BT_TICK(3);
    // ---- RK4 window: stage values ping-pong through shared memory; the per-step bases
    //      xh = x + h/2 F, xd = x + h F and the weighted sum stay in registers (x = analysis now)
    double xh[PPT], xd[PPT];
    const double h2F = h2 * F, dtF = dt * F;
#pragma unroll
    for (int q = 0; q < PPT; ++q) { xh[q] = x[q] + h2F; xd[q] = x[q] + dtF; }
    if (BLOCKED) {
      // the thread's points live in registers; per stage: 3 neighbour loads, PPT stores
      double s[PPT];
#pragma unroll
      for (int q = 0; q < PPT; ++q) s[q] = x[q];
      const int h_l2 = o_m2[0], h_l1 = o_m1[0], h_r = npts > 0 ? o_p1[npts - 1] : 0;
      for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
        for (int st = 0; st < 4; ++st) {
          const double L2 = cur[h_l2], L1 = cur[h_l1], R = cur[h_r];
          double v[PPT];
#pragma unroll
          for (int q = 0; q < PPT; ++q) {
            const double sm2 = q >= 2 ? s[q - 2] : (q == 1 ? L1 : L2);
            const double sm1 = q >= 1 ? s[q - 1] : L1;
            const double sp1 = (q + 1 < PPT && q + 1 < npts) ? s[q + 1 < PPT ? q + 1 : q] : R;
            const double t = fma(sp1 - sm2, sm1, -s[q]);
            if (st == 0) { acc[q] = fma(h6, t, xd[q]); v[q] = fma(h2, t, xh[q]); }
            else if (st == 1) { acc[q] = fma(h3, t, acc[q]); v[q] = fma(h2, t, xh[q]); }
            else if (st == 2) { acc[q] = fma(h3, t, acc[q]); v[q] = fma(dt, t, xd[q]); }
            else { v[q] = fma(h6, t, acc[q]); xh[q] = v[q] + h2F; xd[q] = v[q] + dtF; }
          }
#pragma unroll
          for (int q = 0; q < PPT; ++q) {
            s[q] = v[q];
            if (q < npts) nxt[off[q]] = v[q];
          }
          __syncthreads();
          { double* t_ = cur; cur = nxt; nxt = t_; }
        }
      }
    } else {
    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        // all loads first, all stores last: `cur` and `nxt` are plain pointers into the same array,
        // so a store between the loads would serialise the points of a thread
        double v[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          const int o = off[q] >= 0 ? off[q] : 0;
          const double t = fma(cur[off[q] >= 0 ? o_p1[q] : 0] - cur[off[q] >= 0 ? o_m2[q] : 0],
                               cur[off[q] >= 0 ? o_m1[q] : 0], -cur[o]);
          if (st == 0) { acc[q] = fma(h6, t, xd[q]); v[q] = fma(h2, t, xh[q]); }
          else if (st == 1) { acc[q] = fma(h3, t, acc[q]); v[q] = fma(h2, t, xh[q]); }
          else if (st == 2) { acc[q] = fma(h3, t, acc[q]); v[q] = fma(dt, t, xd[q]); }
          else { v[q] = fma(h6, t, acc[q]); xh[q] = v[q] + h2F; xd[q] = v[q] + dtF; }
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q)
          if (off[q] >= 0) nxt[off[q]] = v[q];
        __syncthreads();
        { double* t_ = cur; cur = nxt; nxt = t_; }
      }
    }
    }
  }
  BT_TICK(4);
#ifdef DAB_BATCH_TIMING
  if (b == 0 && tid == 0)
    printf("batched k=%d N=%d cycles=%d, clocks per cycle: obs %lld stats %lld K4 %lld transform+out %lld rk4 %lld\n", k, N,
           a.n_cycles, bt_[0] / a.n_cycles, bt_[1] / a.n_cycles, bt_[2] / a.n_cycles, bt_[3] / a.n_cycles,
           (bt_[4] + bt_[5]) / a.n_cycles);
#endif
  if (a.out_final != nullptr) {
    double* of = a.out_final + (long long)b * kN;
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (off[q] >= 0) of[gix[q]] = cur[off[q]];
  }
}

// the point mapping the launch will use: blocked with the fewest points per thread that fits 256
// threads, else round-robin
static void batched_mapping(int k, int N, int* ppt, bool* blocked) {
  for (int p = 4; p <= 16; p *= 2)
    if ((long long)k * ((N + p - 1) / p) <= 256) { *ppt = p; *blocked = true; return; }
  const int r = (k * N + 255) / 256;
  *ppt = r <= 4 ? 4 : (r <= 8 ? 8 : 16);
  *blocked = false;
}

size_t batched_smem_bytes(int k, int N, int t_pad, int n_obs) {
  const int nyp = t_pad * n_obs;
  int ppt; bool blocked;
  batched_mapping(k, N, &ppt, &blocked);
  size_t d = 2 * (size_t)batched_state_doubles(k, N, ppt, blocked) + (size_t)k * (nyp | 1) + nyp + ((size_t)k * k + k) + (size_t)k * k +
             (size_t)(eig_small_smem_doubles(k) > ns_small_smem_doubles() ? eig_small_smem_doubles(k)
                                                                          : ns_small_smem_doubles());
  return d * sizeof(double) + 16;
}

cudaError_t batched_launch(const BatchedArgs& a, int batch, cudaStream_t st) {
  const size_t smem = batched_smem_bytes(a.k, a.N, a.t_pad, a.n_obs);
#define DAB_B_CASE(P, B)                                                                              \
  {                                                                                                   \
    cudaFuncSetAttribute(etkf_batched_kernel<P, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    etkf_batched_kernel<P, B><<<batch, 256, smem, st>>>(a);                                           \
    return cudaGetLastError();                                                                        \
  }
  int ppt; bool blocked;
  batched_mapping(a.k, a.N, &ppt, &blocked);
  if (blocked) {
    if (ppt == 4) DAB_B_CASE(4, true)
    if (ppt == 8) DAB_B_CASE(8, true)
    DAB_B_CASE(16, true)
  }
  if (ppt == 4) DAB_B_CASE(4, false)
  if (ppt == 8) DAB_B_CASE(8, false)
  DAB_B_CASE(16, false)
#undef DAB_B_CASE
}

}  // namespace dab
