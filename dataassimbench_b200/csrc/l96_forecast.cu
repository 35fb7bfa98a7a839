// K1 -- Lorenz-96 RK4 ensemble forecast, one analysis window per launch.
//
// Replaces, for the whole ensemble at once, the reference's per-member forecast loop
//   ETKF._step_forecast            dabench/dacycler/_etkf.py:64-80
//   Model.forecast -> Data.generate dabench/data/_data.py:86-211
//   integrate / odeint              dabench/data/_utils.py:16-57
//   Lorenz96.rhs                    dabench/data/_lorenz96.py:119-151
// (fixed-step classical RK4 instead of adaptive Dormand-Prince, as BASELINE.json's north star asks).
//
// Design (B200, FP64-pipe bound: 17 DFMA/DADD per grid point per step, 16 bytes of HBM per
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
  // RK4 coefficients, precomputed on the host so that the kernel takes them from the constant bank
  // (uniform registers): a DFMA with three distinct REGISTER operands issues every 3 cycles (two
  // register-file banks, tools/l96_probe2.cu: 65 % of the FP64 rate), with one uniform operand every 2.
  double h2, h3, h6, h2F, dtF;
  ObsEmit emit;              // optional compact copy of the observed columns (kernels.h)
};

// -DDAB_L96_TRACE (debug builds, tools/build_timing.sh): every CTA records {start, loaded, computed,
// end} (globaltimer ns) and its SM id into the buffer set by dab_debug_l96_trace().
#ifdef DAB_L96_TRACE
__device__ long long* g_l96_trace = nullptr;
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid() { int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define L96_TRACE(i) do { if (g_l96_trace && threadIdx.x == 0) g_l96_trace[8 * blockIdx.x + (i)] = gtime(); } while (0)
#else
#define L96_TRACE(i) do { } while (0)
#endif

template <int P>
__device__ __forceinline__ int phys(int e) { return e + e / P; }   // +1 double per P: conflict-free blocked reads

template <int P, bool PERIODIC, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
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
  L96_TRACE(0);
#ifdef DAB_L96_TRACE
  if (g_l96_trace && threadIdx.x == 0) g_l96_trace[8 * blockIdx.x + 4] = smid();
#endif
  // candidate rows of the compact observation copy (ObsEmit): only the two row bounds are requested here; nothing
  // waits for them until the tile's own loads are in flight, and the columns then travel by cp.async behind the
  // computation (measured: the route costs the kernel 4 us at C3 whether or not this latency is hidden)
  int* s_cols = reinterpret_cast<int*>(stage + phys<P>(T * P));
  int r_lo = 0, r_hi = 0;
  const bool emit = a.emit.yb != nullptr;
  if (emit) {
    long long lo = tile_start, hi = tile_start + a.useful;
    if (hi > a.n) hi = a.n;
    if (hi > lo) { r_lo = __ldg(a.emit.row_start + (lo >> 6)); r_hi = __ldg(a.emit.row_start + ((hi - 1) >> 6) + 1); }
  }
  pdl_wait();                 // the input is the previous kernel's output (transform or forecast chunk)
  pdl_launch_dependents();

  // ---- coalesced load into padded shared memory
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = t + j * T;
    long long g = g0 + e;
    double v;
    if (PERIODIC) {
      if ((unsigned long long)g >= (unsigned long long)a.n) g = pmod(g, a.n);   // only tiles at the seam
      v = in_row[g];
    } else {
      g = g < a.in_lo ? a.in_lo : (g >= a.in_hi ? a.in_hi - 1 : g);
      v = in_row[a.in_base + g];
    }
    stage[phys<P>(e)] = v;
  }
  if (emit) {
    // the columns of the candidate rows travel to shared memory while the tile computes; thread t copies and later
    // reads rows r_lo + t + T i only, so its own cp.async.wait_group is all the ordering the copy needs
    if (r_hi > a.emit.n_rows) r_hi = a.emit.n_rows;
    if (r_hi - r_lo > T * P + 128) r_hi = r_lo + T * P + 128;      // cannot happen for one row per column
    for (int r = r_lo + t; r < r_hi; r += T)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(s_cols + (r - r_lo))),
                   "l"(a.emit.loc + r) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __syncthreads();
  L96_TRACE(1);

  // 17 FP64 instructions per point and step (DADD d, DFMA t, DFMA next stage, DFMA weighted sum;
  // the last stage needs no next-stage value, plus two DADDs per step for the bases):
  //   t_j = (y_{i+1} - y_{i-2}) y_{i-1} - y_i           (k_j = t_j + F: F is folded into the bases)
  //   y_2 = xh + h/2 t_1,  y_3 = xh + h/2 t_2,  y_4 = xd + h t_3,    xh = x + h/2 F,  xd = x + h F
  //   x'  = xd + h/6 t_1 + h/3 t_2 + h/3 t_3 + h/6 t_4  (= x + h/6 (k_1 + 2 k_2 + 2 k_3 + k_4))
  double s[P], xh[P], xd[P], acc[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    s[i] = stage[phys<P>(t * P + i)];
    xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF;
  }

  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;

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
      // next-stage coefficient ca and RK4 weight cw: uniform operands (see L96Args)
#define L96_CA (st == 2 ? a.dt : a.h2)
#define L96_CW ((st == 0 || st == 3) ? a.h6 : a.h3)
      double kk[P];
      // edge points first so the next stage's edges can be published early
      kk[0] = fma((P > 1 ? s[1] : R) - L.x, L.y, -s[0]);
      if (P > 1) kk[1] = fma((P > 2 ? s[2] : R) - L.y, s[0], -s[1]);
      if (P > 2) kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
      if (P > 3) kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
      if (st < 3) {
        const double n0 = fma(L96_CA, kk[0], st == 2 ? xd[0] : xh[0]);
        const double n2 = fma(L96_CA, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
        const double n1 = fma(L96_CA, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
        e_first[buf * T + t] = n0;
        e_last2[buf * T + t] = make_double2(n2, n1);
      } else {
        const double n0 = fma(L96_CW, kk[0], acc[0]);
        const double n2 = fma(L96_CW, kk[P - 2], acc[P - 2]);
        const double n1 = fma(L96_CW, kk[P - 1], acc[P - 1]);
        e_first[buf * T + t] = n0;
        e_last2[buf * T + t] = make_double2(n2, n1);
      }
#pragma unroll
      for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (st == 0) acc[i] = fma(L96_CW, kk[i], xd[i]);
        else if (st < 3) acc[i] = fma(L96_CW, kk[i], acc[i]);
        if (st < 3) s[i] = fma(L96_CA, kk[i], st == 2 ? xd[i] : xh[i]);
        else { s[i] = fma(L96_CW, kk[i], acc[i]); xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
      }
#undef L96_CA
#undef L96_CW
    }
    if (a.traj != nullptr) {
      double* trow = a.traj + (long long)step * a.traj_stride + (long long)member * a.out_ld;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const long long g = g0 + (long long)t * P + i;
        if (g >= out_lo && g < out_hi) trow[g] = s[i];
      }
    }
  }

  // ---- coalesced store of the useful (still valid) centre of the tile
  __syncthreads();
  L96_TRACE(2);
#pragma unroll
  for (int i = 0; i < P; ++i) stage[phys<P>(t * P + i)] = s[i];
  __syncthreads();
  double* out_row = a.out + (long long)member * a.out_ld;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = t + j * T;
    const long long g = g0 + e;
    if (g >= out_lo && g < out_hi) out_row[g] = stage[phys<P>(e)];
  }
  if (emit) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    double* yb_row = a.emit.yb + (long long)member * a.emit.ldy;
    for (int r = r_lo + t; r < r_hi; r += T) {
      const long long c = s_cols[r - r_lo];
      if (c >= out_lo && c < out_hi) yb_row[r] = stage[phys<P>((int)(c - g0))];
    }
  }
  L96_TRACE(3);
}

#ifdef DAB_L96_TRACE
}  // namespace dab
extern "C" __attribute__((visibility("default"))) int dab_debug_l96_trace(long long* dev_buf) {
  return (int)cudaMemcpyToSymbol(dab::g_l96_trace, &dev_buf, sizeof(dev_buf));
}
namespace dab {
#endif

// ------------------------------------------------------------------------------------------
// Host-side launch planning.  A launch advances at most `max_steps_per_launch` steps (halo =
// 12 points per step must stay a modest fraction of the tile); longer windows are chunked.
// + the tile's candidate observation columns (ObsEmit): at most W + 128 ints, fetched while the tile loads
static size_t l96_smem_bytes(int T, int P, bool emit) {
  const int W = T * P;
  return ((size_t)(W + W / P) + 6 * (size_t)T) * sizeof(double) + (emit ? (size_t)(W + 128) * sizeof(int) : 0);
}

int l96_max_steps_per_launch() { return 32; }

template <int P, int MAXT, int MINB>
static cudaError_t launch_pb(const L96Args& a, int k, int T, bool periodic, cudaStream_t st) {
  const size_t smem = l96_smem_bytes(T, P, a.emit.yb != nullptr);
  dim3 grid((unsigned)((long long)a.tiles_per_member * k));
  if (periodic) {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, true, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return launch_chain(true, l96_rk4_window_kernel<P, true, MAXT, MINB>, grid, dim3(T), smem, st, a);
  } else {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, false, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return launch_chain(true, l96_rk4_window_kernel<P, false, MAXT, MINB>, grid, dim3(T), smem, st, a);
  }
}

// CTAs of <= 256 threads use the register-capped build (<= 80 registers: three CTAs per SM, so
// that one CTA's barrier/exchange bubbles are covered by the other two); wider CTAs the plain one.
// Kernel builds.  8 points per thread: <= 80 registers, 3 (T <= 256) or 2 (T <= 384) CTAs per SM, or
// the uncapped build for wider CTAs.  16 points per thread (half the exchanges per flop, half the
// halo share at the same thread count): <= 255 registers (two 128-thread CTAs per SM), or the
// 170-register build `dense` (three 128-thread or two 192-thread CTAs per SM).  12 points per thread: <= 168
// registers, two 192-thread CTAs per SM.  The autotuner (engine.cu) measures them per shape.
template <int P>
static cudaError_t launch_p(const L96Args& a, int k, int T, bool periodic, bool dense, cudaStream_t st) {
  if constexpr (P == 8) {
    if (T <= 256) return launch_pb<P, 256, 3>(a, k, T, periodic, st);
    if (T <= 384) return launch_pb<P, 384, 2>(a, k, T, periodic, st);
    return launch_pb<P, 512, 1>(a, k, T, periodic, st);
  } else if constexpr (P == 16) {
    if (dense && T <= 128) return launch_pb<P, 128, 3>(a, k, T, periodic, st);
    if (dense && T <= 192) return launch_pb<P, 192, 2>(a, k, T, periodic, st);
    return launch_pb<P, 256, 1>(a, k, T, periodic, st);
  } else if constexpr (P == 12) {
    return launch_pb<P, 192, 2>(a, k, T <= 192 ? T : 192, periodic, st);
  } else {
    return launch_pb<P, 512, 1>(a, k, T, periodic, st);
  }
}

// CTA width for one launch of `steps` steps over slabs of n points and k members when no shape is
// forced: 8 points per thread, 256 threads (three CTAs per SM), narrower for small problems so that
// more SMs get a tile, wider when the halo would exceed half of the tile.  Problems large enough to
// matter do not come through here: the cycler measures all shapes at setup (engine.cu) and passes
// the winner as `forced_T`.
static bool l96_plan(long long n, int k, int steps, int sms, int P, int t_max, int forced_T,
                     int* T_out, int* useful_out, int* tiles_out) {
  const int halo = 12 * steps;
  int T = 256;
  if (forced_T >= 32 && forced_T <= 512 && forced_T % 32 == 0 && forced_T * P >= halo + 16) {
    T = forced_T;
  } else {
    while (T * P < 2 * halo + 16 && T < 512) T += 32;      // keep the halo below half of the tile
    while (T > 64 && ceil_div_ll(n, T * P - halo) * k < 3ll * sms && (T / 2) * P >= 2 * halo + 16) T /= 2;
  }
  // the launch is bounded by the build's __launch_bounds__: plan for the width that will really run
  if (T > t_max) T = t_max;
  if (T * P < halo + 16) return false;                     // this build cannot hold the window's halo
  const long long tiles = ceil_div_ll(n, T * P - halo);
  *T_out = T;
  *tiles_out = (int)tiles;
  *useful_out = (int)ceil_div_ll(n, tiles);                // balance the tiles
  return true;
}

cudaError_t l96_forecast_launch(const double* in, double* out, double* traj, long long n, int k,
                                int n_steps, double dt, double F, long long in_ld, long long out_ld,
                                bool periodic, long long in_base, long long in_lo, long long in_hi,
                                long long traj_stride, int sms, int forced_T, cudaStream_t st, const ObsEmit* emit) {
  // CTA shape code: plain T = 8 points per thread; 1000 * c + T with c = 12 / 16 points per thread,
  // c = 17: 16 points per thread in the 170-register build (kernels.h: l96_shape_code).
  if (const char* ev = getenv("DAB_L96_T")) {            // experiment/test knob: force the CTA shape
    const int t = atoi(ev) % 1000;
    if (t >= 32 && t <= 512 && t % 32 == 0) forced_T = atoi(ev);
  }
  int P = 8;
  bool dense = false;
  if (forced_T >= 1000) {
    const int c = forced_T / 1000;
    forced_T %= 1000;
    P = c == 12 ? 12 : 16;
    dense = c == 17;
  }
  int T = 256, useful = 0, tiles = 0;
  int tmax = P == 8 ? 512 : (P == 12 ? 192 : (dense ? 192 : 256));
  if (forced_T > tmax) forced_T = tmax;
  if (!l96_plan(n, k, n_steps, sms, P, tmax, forced_T, &T, &useful, &tiles)) {
    // a forced 12/16-point shape too narrow for this window's halo: the 8-point build (up to 512 threads) holds it
    P = 8; dense = false; tmax = 512;
    if (!l96_plan(n, k, n_steps, sms, P, tmax, 0, &T, &useful, &tiles)) return cudaErrorInvalidValue;
  }
  L96Args a;
  a.in = in; a.out = out; a.traj = traj;
  a.in_ld = in_ld; a.out_ld = out_ld; a.in_base = in_base; a.in_lo = in_lo; a.in_hi = in_hi;
  a.n = n; a.traj_stride = traj_stride;
  a.n_steps = n_steps; a.halo_l = 8 * n_steps; a.useful = useful; a.tiles_per_member = tiles;
  a.dt = dt; a.F = F;
  a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  if (emit != nullptr) a.emit = *emit;
  if (P == 12) return launch_p<12>(a, k, T, periodic, dense, st);
  if (P == 16) return launch_p<16>(a, k, T, periodic, dense, st);
  return launch_p<8>(a, k, T, periodic, dense, st);
}

}  // namespace dab
