// K1 -- Lorenz-96 RK4 ensemble forecast, one analysis window per launch.
//
// Replaces, for the whole ensemble at once, the reference's per-member forecast loop
//   ETKF._step_forecast            dabench/dacycler/_etkf.py:64-80
//   Model.forecast -> Data.generate dabench/data/_data.py:86-211
//   integrate / odeint              dabench/data/_utils.py:16-57
//   Lorenz96.rhs                    dabench/data/_lorenz96.py:119-151
// (fixed-step classical RK4 instead of adaptive Dormand-Prince, as BASELINE.json's north star asks).
//
// Design (B200, FP64-pipe bound: 19 DFMA/DADD per grid point per step, 16 bytes of HBM per
// grid point per WINDOW):
//   * one CTA = one member x one tile of W = blockDim*P consecutive grid points, held in
//     REGISTERS (P points per thread, blocked) for all n_steps steps of the window;
//   * overlapped temporal tiling: the tile carries a halo of 8*n_steps points on the left and
//     4*n_steps on the right (the stencil reaches i-2, i-1, i+1 => 2 / 1 per RK stage); the halo
//     is recomputed redundantly and goes stale from the outside in, so CTAs never talk to each
//     other and HBM is touched only at window boundaries;
//   * per RK stage a thread needs 3 neighbour values (2 from the left thread, 1 from the right):
//     they go through a double-buffered shared-memory edge array, one __syncthreads per stage;
//   * loads/stores are staged through padded shared memory so HBM sees fully coalesced 8-byte
//     accesses while each thread ends up with P consecutive points.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dab {

struct L96Args {
  const double* in;
  double* out;
  double* traj;              // nullable; [n_steps][k][out_ld]: state after step s+1 at traj + s*traj_stride
  long long in_ld, out_ld;   // row (member) strides in doubles
  long long in_base;         // EXT mode: column of slab position 0 inside an input row
  long long in_lo, in_hi;    // EXT mode: valid slab-coordinate range [in_lo, in_hi) of an input row
  long long n;               // PERIODIC: ring size; EXT: slab length
  long long traj_stride;     // doubles between consecutive trajectory steps
  int n_steps, halo_l, useful, tiles_per_member;
  double dt, F;
};

template <int P>
__device__ __forceinline__ int phys(int e) { return e + e / P; }   // +1 double per P: conflict-free blocked reads

template <int P, bool PERIODIC>
__global__ void __launch_bounds__(512)
l96_rk4_window_kernel(const L96Args a) {
  extern __shared__ __align__(16) double smem[];
  const int T = blockDim.x;
  const int t = threadIdx.x;
  double2* e_last2 = reinterpret_cast<double2*>(smem);  // [2][T] (s_{P-2}, s_{P-1}) of every thread
  double* e_first = smem + 4 * T;                       // [2][T] s_0 of every thread
  double* stage = e_first + 2 * T;                      // phys(W) doubles, load/store staging

  const int member = blockIdx.x / a.tiles_per_member;
  const int tile = blockIdx.x - member * a.tiles_per_member;
  const long long tile_start = (long long)tile * a.useful;   // slab coordinate of first useful point
  const long long g0 = tile_start - a.halo_l;                // slab coordinate of ext position 0
  const double* in_row = a.in + (long long)member * a.in_ld;

  // ---- coalesced load into padded shared memory
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = t + j * T;
    long long g = g0 + e;
    double v;
    if (PERIODIC) {
      v = in_row[pmod(g, a.n)];
    } else {
      g = g < a.in_lo ? a.in_lo : (g >= a.in_hi ? a.in_hi - 1 : g);
      v = in_row[a.in_base + g];
    }
    stage[phys<P>(e)] = v;
  }
  __syncthreads();

  double x[P], s[P], acc[P];
#pragma unroll
  for (int i = 0; i < P; ++i) { x[i] = stage[phys<P>(t * P + i)]; s[i] = x[i]; }

  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;
  const double dt = a.dt, F = a.F;
  const double h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0;

  // edges of the stage-1 input (x itself)
  int buf = 0;
  e_first[t] = s[0];
  e_last2[t] = make_double2(s[P - 2], s[P - 1]);

  const long long out_lo = tile_start;
  long long out_hi = tile_start + a.useful;
  if (out_hi > a.n) out_hi = a.n;

  for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      __syncthreads();
      const double2 L = e_last2[buf * T + tl];
      const double R = e_first[buf * T + tr];
      buf ^= 1;
      const double ca = (st == 2) ? dt : h2;                 // next-stage coefficient
      const double cw = (st == 0 || st == 3) ? h6 : h3;      // RK4 weight
      double kk[P];
      // edge points first so the next stage's edges can be published early
      kk[0] = fma((P > 1 ? s[1] : R) - L.x, L.y, F - s[0]);
      if (P > 1) kk[1] = fma((P > 2 ? s[2] : R) - L.y, s[0], F - s[1]);
      if (P > 2) kk[P - 1] = fma(R - s[P - 3], s[P - 2], F - s[P - 1]);
      if (P > 3) kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], F - s[P - 2]);
      if (st < 3) {
        const double n0 = fma(ca, kk[0], x[0]);
        const double n2 = fma(ca, kk[P - 2], x[P - 2]);
        const double n1 = fma(ca, kk[P - 1], x[P - 1]);
        e_first[buf * T + t] = n0;
        e_last2[buf * T + t] = make_double2(n2, n1);
      } else {
        const double n0 = fma(cw, kk[0], acc[0]);
        const double n2 = fma(cw, kk[P - 2], acc[P - 2]);
        const double n1 = fma(cw, kk[P - 1], acc[P - 1]);
        e_first[buf * T + t] = n0;
        e_last2[buf * T + t] = make_double2(n2, n1);
      }
#pragma unroll
      for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], F - s[i]);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (st == 0) acc[i] = fma(cw, kk[i], x[i]);
        else acc[i] = fma(cw, kk[i], acc[i]);
        if (st < 3) s[i] = fma(ca, kk[i], x[i]);
        else { x[i] = acc[i]; s[i] = acc[i]; }
      }
    }
    if (a.traj != nullptr) {
      double* trow = a.traj + (long long)step * a.traj_stride + (long long)member * a.out_ld;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const long long g = g0 + (long long)t * P + i;
        if (g >= out_lo && g < out_hi) trow[g] = x[i];
      }
    }
  }

  // ---- coalesced store of the useful (still valid) centre of the tile
  __syncthreads();
#pragma unroll
  for (int i = 0; i < P; ++i) stage[phys<P>(t * P + i)] = x[i];
  __syncthreads();
  double* out_row = a.out + (long long)member * a.out_ld;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = t + j * T;
    const long long g = g0 + e;
    if (g >= out_lo && g < out_hi) out_row[g] = stage[phys<P>(e)];
  }
}

// ------------------------------------------------------------------------------------------
// Host-side launch planning.  A launch advances at most `max_steps_per_launch` steps (halo =
// 12 points per step must stay a modest fraction of the tile); longer windows are chunked.
static size_t l96_smem_bytes(int T, int P) {
  const int W = T * P;
  return ((size_t)(W + W / P) + 6 * (size_t)T) * sizeof(double);
}

int l96_max_steps_per_launch() { return 32; }

template <int P>
static cudaError_t launch_p(const L96Args& a, int k, int T, bool periodic, cudaStream_t st) {
  const size_t smem = l96_smem_bytes(T, P);
  dim3 grid((unsigned)((long long)a.tiles_per_member * k));
  if (periodic) {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_kernel<P, true><<<grid, T, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_kernel<P, false><<<grid, T, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// Pick the CTA width T for one launch of `steps` steps over slabs of n points and k members on
// `sms` SMs.  The kernel is FP64-pipe bound, so its time is proportional to
//   waves * (points computed per SM per wave) = ceil(ctas / (sms*occ)) * occ * W,
// halo recomputation and wave quantisation included; ties go to more resident threads.
static void l96_plan(long long n, int k, int steps, int sms, int P, int regs_per_thread, int forced_T,
                     int* T_out, int* useful_out, int* tiles_out) {
  const int halo = 12 * steps;
  double best = 1e300;
  int t_lo = 64, t_hi = 512;
  if (const char* ev = getenv("DAB_L96_T")) {            // experiment knob: force the CTA width
    const int t = atoi(ev);
    if (t >= 32 && t <= 512 && t % 32 == 0) forced_T = t;
  }
  if (forced_T >= 32 && forced_T <= 512 && forced_T % 32 == 0 && forced_T * P >= halo + 16) {
    t_lo = forced_T; t_hi = forced_T;
  }
  for (int T = t_lo; T <= t_hi; T += 32) {
    const int W = T * P;
    if (W < halo + 16) continue;
    const int useful_max = W - halo;
    const long long tiles = ceil_div_ll(n, useful_max);
    int occ = 65536 / (regs_per_thread * T);
    const size_t smem = l96_smem_bytes(T, P) + 1024;
    if (occ > (int)(227 * 1024 / smem)) occ = (int)(227 * 1024 / smem);
    if (occ > 2048 / T) occ = 2048 / T;
    if (occ < 1) occ = 1;
    const long long ctas = tiles * k;
    if (ctas < (long long)sms * occ) occ = (int)ceil_div_ll(ctas, sms);   // small problems: spread out
    const long long waves = ceil_div_ll(ctas, (long long)sms * occ);
    double cost = (double)waves * occ * W;
    if (occ * T < 512) cost *= 512.0 / (occ * T);       // too few warps to cover DFMA/barrier latency
    cost *= 1.0 - 1e-4 * (occ * T) / 2048.0;
    if (cost < best) {
      best = cost; *T_out = T; *tiles_out = (int)tiles;
      *useful_out = (int)ceil_div_ll(n, tiles);           // balance the tiles
    }
  }
}

cudaError_t l96_forecast_launch(const double* in, double* out, double* traj, long long n, int k,
                                int n_steps, double dt, double F, long long in_ld, long long out_ld,
                                bool periodic, long long in_base, long long in_lo, long long in_hi,
                                long long traj_stride, int sms, int forced_T, cudaStream_t st) {
  int P = 8;
  if (const char* ev = getenv("DAB_L96_P")) { if (atoi(ev) == 4) P = 4; }   // experiment knob
  int T = 256, useful = 0, tiles = 0;
  l96_plan(n, k, n_steps, sms, P, P == 8 ? 88 : 72, forced_T, &T, &useful, &tiles);
  L96Args a;
  a.in = in; a.out = out; a.traj = traj;
  a.in_ld = in_ld; a.out_ld = out_ld; a.in_base = in_base; a.in_lo = in_lo; a.in_hi = in_hi;
  a.n = n; a.traj_stride = traj_stride;
  a.n_steps = n_steps; a.halo_l = 8 * n_steps; a.useful = useful; a.tiles_per_member = tiles;
  a.dt = dt; a.F = F;
  if (P == 4) return launch_p<4>(a, k, T, periodic, st);
  return launch_p<8>(a, k, T, periodic, st);
}

}  // namespace dab
