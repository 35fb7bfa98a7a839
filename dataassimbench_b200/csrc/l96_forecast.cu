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
};

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

  // 17 FP64 instructions per point and step (DADD d, DFMA t, DFMA next stage, DFMA weighted sum;
  // the last stage needs no next-stage value, plus two DADDs per step for the bases):
  //   t_j = (y_{i+1} - y_{i-2}) y_{i-1} - y_i           (k_j = t_j + F: F is folded into the bases)
  //   y_2 = xh + h/2 t_1,  y_3 = xh + h/2 t_2,  y_4 = xd + h t_3,    xh = x + h/2 F,  xd = x + h F
  //   x'  = xd + h/6 t_1 + h/3 t_2 + h/3 t_3 + h/6 t_4  (= x + h/6 (k_1 + 2 k_2 + 2 k_3 + k_4))
  double s[P], xh[P], xd[P], acc[P];
  const double dt = a.dt, F = a.F;
  const double h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0;
  const double h2F = h2 * F, dtF = dt * F;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    s[i] = stage[phys<P>(t * P + i)];
    xh[i] = s[i] + h2F; xd[i] = s[i] + dtF;
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
      const double ca = (st == 2) ? dt : h2;                 // next-stage coefficient
      const double cw = (st == 0 || st == 3) ? h6 : h3;      // RK4 weight
      double kk[P];
      // edge points first so the next stage's edges can be published early
      kk[0] = fma((P > 1 ? s[1] : R) - L.x, L.y, -s[0]);
      if (P > 1) kk[1] = fma((P > 2 ? s[2] : R) - L.y, s[0], -s[1]);
      if (P > 2) kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
      if (P > 3) kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
      if (st < 3) {
        const double n0 = fma(ca, kk[0], st == 2 ? xd[0] : xh[0]);
        const double n2 = fma(ca, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
        const double n1 = fma(ca, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
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
      for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (st == 0) acc[i] = fma(cw, kk[i], xd[i]);
        else if (st < 3) acc[i] = fma(cw, kk[i], acc[i]);
        if (st < 3) s[i] = fma(ca, kk[i], st == 2 ? xd[i] : xh[i]);
        else { s[i] = fma(cw, kk[i], acc[i]); xh[i] = s[i] + h2F; xd[i] = s[i] + dtF; }
      }
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
}

// ------------------------------------------------------------------------------------------
// v2: persistent CTAs, cp.async prefetch of the next tile, shuffle-based neighbour exchange.
//
// ncu on v1 (profiles/): FP64 pipe 62 % active; warps stalled on the tile load at CTA start
// (long scoreboard 14.5 %), on the per-stage shared-memory exchange (short scoreboard 10.8 %)
// and on the four barriers per step (5.9 %).  v2 removes all three:
//   * a warp owns 256 consecutive points (8 per lane); per RK stage a lane gets its three
//     neighbour values with warp shuffles -- no shared memory, no barrier inside a step;
//   * the warp tile carries its own 8-left / 4-right halo for ONE step (2 / 1 points go stale
//     per stage), so warps exchange 12 points once per step through a double-buffered shared
//     array: one __syncthreads per step instead of four, 244 of 256 points useful;
//   * CTAs are persistent and stream tiles; the next tile is fetched with cp.async
//     (global -> shared, no registers) while the current one is integrated.
constexpr int WU = 244;   // useful points of a 256-point warp tile (8 + 4 per-step halo)

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

template <bool PERIODIC>
__global__ void __launch_bounds__(512)
l96_rk4_window_v2_kernel(const L96Args a, int total_tiles) {
  constexpr int P = 8;
  extern __shared__ __align__(16) double smem[];
  const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5, NW = T >> 5;
  const int Wc = NW * WU + 12;                          // points of the CTA tile
  const int SW = phys<P>(Wc) + 2;                       // staging buffer length (doubles)
  double* toRight = smem;                               // [2][NW][8]  points 244..251 of every warp
  double* toLeft = smem + 16 * NW;                      // [2][NW][4]  points 8..11 of every warp
  double* stage = smem + 24 * NW;                       // [2][SW]

  const double dt = a.dt, F = a.F;
  const double h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0;
  const int eb = w * WU + P * lane;                     // CTA-tile position of this lane's first point

  auto issue_load = [&](int tile_id, double* st) {
    const int member = tile_id / a.tiles_per_member;
    const int tile = tile_id - member * a.tiles_per_member;
    const long long g0 = (long long)tile * a.useful - a.halo_l;
    const double* in_row = a.in + (long long)member * a.in_ld;
    for (int e = t; e < Wc; e += T) {
      long long g = g0 + e;
      const double* src;
      if (PERIODIC) {
        src = in_row + pmod(g, a.n);
      } else {
        g = g < a.in_lo ? a.in_lo : (g >= a.in_hi ? a.in_hi - 1 : g);
        src = in_row + a.in_base + g;
      }
      cp_async8(&st[phys<P>(e)], src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  int tile_id = blockIdx.x;
  if (tile_id < total_tiles) issue_load(tile_id, stage);
  for (; tile_id < total_tiles; tile_id += gridDim.x, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double* st_cur = stage + buf * SW;
    double x[P], s[P], acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) { x[i] = st_cur[phys<P>(eb + i)]; s[i] = x[i]; }
    {
      const int next = tile_id + gridDim.x;
      if (next < total_tiles) issue_load(next, stage + (buf ^ 1) * SW);
    }
    const int member = tile_id / a.tiles_per_member;
    const int tile = tile_id - member * a.tiles_per_member;
    const long long tile_start = (long long)tile * a.useful;
    const long long g0 = tile_start - a.halo_l;
    const long long out_lo = tile_start;
    long long out_hi = tile_start + a.useful;
    if (out_hi > a.n) out_hi = a.n;

    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
      for (int stg = 0; stg < 4; ++stg) {
        const double Lx = __shfl_up_sync(0xffffffffu, s[P - 2], 1);
        const double Ly = __shfl_up_sync(0xffffffffu, s[P - 1], 1);
        const double R = __shfl_down_sync(0xffffffffu, s[0], 1);
        const double ca = (stg == 2) ? dt : h2;
        const double cw = (stg == 0 || stg == 3) ? h6 : h3;
        double kk[P];
#pragma unroll
        for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], F - s[i]);
        kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], F - s[P - 2]);
        kk[0] = fma(s[1] - Lx, Ly, F - s[0]);
        kk[1] = fma(s[2] - Ly, s[0], F - s[1]);
        kk[P - 1] = fma(R - s[P - 3], s[P - 2], F - s[P - 1]);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (stg == 0) acc[i] = fma(cw, kk[i], x[i]);
          else acc[i] = fma(cw, kk[i], acc[i]);
          if (stg < 3) s[i] = fma(ca, kk[i], x[i]);
          else { x[i] = acc[i]; s[i] = acc[i]; }
        }
      }
      if (a.traj != nullptr) {
        double* trow = a.traj + (long long)step * a.traj_stride + (long long)member * a.out_ld;
#pragma unroll
        for (int i = 0; i < P; ++i) {
          const int loc = P * lane + i;
          const long long g = g0 + eb + i;
          if (loc >= 8 && loc < 8 + WU && g >= out_lo && g < out_hi) trow[g] = x[i];
        }
      }
      if (step + 1 < a.n_steps) {                       // refresh the warp halos for the next step
        const int pb = step & 1;
        double* tr = toRight + (pb * NW + w) * 8;
        double* tl = toLeft + (pb * NW + w) * 4;
        if (lane == 30) { tr[0] = x[4]; tr[1] = x[5]; tr[2] = x[6]; tr[3] = x[7]; }
        if (lane == 31) { tr[4] = x[0]; tr[5] = x[1]; tr[6] = x[2]; tr[7] = x[3]; }
        if (lane == 1) { tl[0] = x[0]; tl[1] = x[1]; tl[2] = x[2]; tl[3] = x[3]; }
        __syncthreads();
        if (lane == 0 && w > 0) {
          const double* src = toRight + (pb * NW + w - 1) * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) { x[i] = src[i]; s[i] = x[i]; }
        }
        if (lane == 31 && w < NW - 1) {
          const double* src = toLeft + (pb * NW + w + 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) { x[4 + i] = src[i]; s[4 + i] = x[4 + i]; }
        }
      }
    }

    // ---- coalesced store of the useful centre (each point written by the warp that owns it)
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const int loc = P * lane + i;
      if (loc >= 8 && loc < 8 + WU) st_cur[phys<P>(eb + i)] = x[i];
    }
    __syncthreads();
    double* out_row = a.out + (long long)member * a.out_ld;
    for (int e = t; e < Wc; e += T) {
      const long long g = g0 + e;
      if (e >= 8 && e < Wc - 4 && g >= out_lo && g < out_hi) out_row[g] = st_cur[phys<P>(e)];
    }
  }
}

static size_t l96_v2_smem_bytes(int T) {
  const int NW = T / 32, Wc = NW * WU + 12;
  return ((size_t)24 * NW + 2 * (size_t)(Wc + Wc / 8 + 2)) * sizeof(double);
}

// v2 launch geometry: CTA width T (multiple of 32), persistent grid.
static cudaError_t l96_v2_launch(L96Args a, long long n, int k, int steps, int T, bool periodic, int sms,
                                 cudaStream_t st) {
  const int NW = T / 32, Wc = NW * WU + 12;
  const int useful_max = Wc - 12 * steps;
  const long long tiles = ceil_div_ll(n, useful_max);
  a.tiles_per_member = (int)tiles;
  a.useful = (int)ceil_div_ll(n, tiles);
  a.halo_l = 8 * steps;
  const long long total = tiles * k;
  const size_t smem = l96_v2_smem_bytes(T);
  int occ = 65536 / (128 * T);
  if (occ > (int)((227 * 1024) / (smem + 1024))) occ = (int)((227 * 1024) / (smem + 1024));
  if (occ < 1) occ = 1;
  long long grid = (long long)sms * occ;
  if (grid > total) grid = total;
  // even out the tiles per CTA: with r = ceil(total/grid) rounds, ceil(total/r) CTAs suffice
  const long long rounds = ceil_div_ll(total, grid);
  grid = ceil_div_ll(total, rounds);
  if (periodic) {
    cudaFuncSetAttribute(l96_rk4_window_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_v2_kernel<true><<<(unsigned)grid, T, smem, st>>>(a, (int)total);
  } else {
    cudaFuncSetAttribute(l96_rk4_window_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_v2_kernel<false><<<(unsigned)grid, T, smem, st>>>(a, (int)total);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Host-side launch planning.  A launch advances at most `max_steps_per_launch` steps (halo =
// 12 points per step must stay a modest fraction of the tile); longer windows are chunked.
static size_t l96_smem_bytes(int T, int P) {
  const int W = T * P;
  return ((size_t)(W + W / P) + 6 * (size_t)T) * sizeof(double);
}

int l96_max_steps_per_launch() { return 32; }

template <int P, int MAXT, int MINB>
static cudaError_t launch_pb(const L96Args& a, int k, int T, bool periodic, cudaStream_t st) {
  const size_t smem = l96_smem_bytes(T, P);
  dim3 grid((unsigned)((long long)a.tiles_per_member * k));
  if (periodic) {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, true, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_kernel<P, true, MAXT, MINB><<<grid, T, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(l96_rk4_window_kernel<P, false, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    l96_rk4_window_kernel<P, false, MAXT, MINB><<<grid, T, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// CTAs of <= 256 threads use the register-capped build (<= 80 registers: three CTAs per SM, so
// that one CTA's barrier/exchange bubbles are covered by the other two); wider CTAs the plain one.
template <int P>
static cudaError_t launch_p(const L96Args& a, int k, int T, bool periodic, cudaStream_t st) {
  if (getenv("DAB_L96_OCC4") && T <= 256 && P == 8) return launch_pb<P, 256, 4>(a, k, T, periodic, st);
  if (T <= 256 && P == 8) return launch_pb<P, 256, 3>(a, k, T, periodic, st);
  if (T <= 384 && P == 8) return launch_pb<P, 384, 2>(a, k, T, periodic, st);
  return launch_pb<P, 512, 1>(a, k, T, periodic, st);
}

// CTA width for one launch of `steps` steps over slabs of n points and k members.  Measured on
// B200 (tools/tune_forecast.py, C3 shape): 256 threads x 3 resident CTAs per SM (80 registers)
// beats every other width by 8-25 % -- three independent CTAs cover each other's per-stage
// exchange bubbles -- so that is the default; small problems use narrower CTAs to reach more SMs.
// The cycler re-checks the choice with CUDA events at setup (engine.cu) and passes `forced_T`.
static void l96_plan(long long n, int k, int steps, int sms, int P, int regs_per_thread, int forced_T,
                     int* T_out, int* useful_out, int* tiles_out) {
  (void)regs_per_thread;
  const int halo = 12 * steps;
  if (const char* ev = getenv("DAB_L96_T")) {            // experiment knob: force the CTA width
    const int t = atoi(ev);
    if (t >= 32 && t <= 512 && t % 32 == 0) forced_T = t;
  }
  int T = 256;
  if (forced_T >= 32 && forced_T <= 512 && forced_T % 32 == 0 && forced_T * P >= halo + 16) {
    T = forced_T;
  } else {
    while (T * P < 2 * halo + 16 && T < 512) T += 32;      // keep the halo below half of the tile
    while (T > 64 && ceil_div_ll(n, T * P - halo) * k < 3ll * sms && (T / 2) * P >= 2 * halo + 16) T /= 2;
  }
  const long long tiles = ceil_div_ll(n, T * P - halo);
  *T_out = T;
  *tiles_out = (int)tiles;
  *useful_out = (int)ceil_div_ll(n, tiles);                // balance the tiles
}

cudaError_t l96_forecast_launch(const double* in, double* out, double* traj, long long n, int k,
                                int n_steps, double dt, double F, long long in_ld, long long out_ld,
                                bool periodic, long long in_base, long long in_lo, long long in_hi,
                                long long traj_stride, int sms, int forced_T, cudaStream_t st) {
  if (const char* ev = getenv("DAB_L96_V2")) {            // experiment knob: force the v2 kernel
    const int t = atoi(ev);
    if (t >= 64 && t <= 512 && t % 32 == 0) forced_T = -t;
  }
  if (forced_T < 0) {                                     // v2 (persistent / shuffle) with CTA width -forced_T
    const int T2 = -forced_T;
    if ((T2 / 32) * WU + 12 > 12 * n_steps + 64) {
      L96Args a2;
      a2.in = in; a2.out = out; a2.traj = traj;
      a2.in_ld = in_ld; a2.out_ld = out_ld; a2.in_base = in_base; a2.in_lo = in_lo; a2.in_hi = in_hi;
      a2.n = n; a2.traj_stride = traj_stride; a2.n_steps = n_steps; a2.dt = dt; a2.F = F;
      a2.halo_l = 0; a2.useful = 0; a2.tiles_per_member = 0;
      return l96_v2_launch(a2, n, k, n_steps, T2, periodic, sms, st);
    }
    forced_T = 0;
  }
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
