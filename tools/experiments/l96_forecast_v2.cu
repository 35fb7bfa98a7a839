// EXPERIMENT (not part of the product build; tools/l96_v2_test.cu runs it): K1 as a persistent dataflow of
// WARP-private windows.  Correct (2e-15 against CPU RK4) but slower than the CTA-tile kernels: 21.3 TFLOP/s at
// C4 against 22.6 for csrc/l96_forecast_v3.cu -- a warp spends ~40 % of its time on per-job bookkeeping (claim,
// flags, gpu-scope fence, staging) that three warps per scheduler do not hide.  profiles/forecast_probes_r02.md.
//
// Replaces the same reference functions as l96_forecast.cu (ETKF._step_forecast
// dabench/dacycler/_etkf.py:64-80 -> Model.forecast -> Data.generate dabench/data/_data.py:86-211 ->
// integrate dabench/data/_utils.py:16-57 with Lorenz96.rhs dabench/data/_lorenz96.py:119-151), fixed-step
// classical RK4 for all members at once.
//
// Why a second design (measurements: tools/l96_probe2.cu, profiles/forecast_probes_r02.md):
//   * the FP64 pipe runs the 17-instruction stage arithmetic at 88 % when warps never wait for each
//     other (neighbour values by warp shuffle), and at 70-80 % as soon as warps meet once per RK stage
//     or step (CTA barrier, named barriers, neighbour flags, mbarriers: all measured);
//   * a CTA-wide tile of 2048 points pays 240 halo points (11.7 %) for a 20-step window.
// So here a job is ONE WARP advancing ONE window of 32 x 16 = 512 consecutive points of one member by a
// CHUNK of <= 6 steps (halo 12 points per step of the chunk: 60 of 512 at 5 steps), entirely in
// registers, with no synchronisation of any kind inside the chunk.  A window of S steps is ceil(S/5)
// chunks; chunk c reads what chunk c-1 wrote (two ping-pong scratch buffers, L2-resident for C3).
// All chunks run in ONE persistent launch: jobs are numbered chunk-major and handed out in that order
// by an atomic dispenser; a job of chunk c waits (per-job flags in global memory) for the <= 3 jobs of
// chunk c-1 whose output it reads or whose input it overwrites.  Those have a lower number, so they
// were claimed earlier by warps that are running: the wait cannot deadlock, and for the bench shapes
// the producers finished thousands of jobs earlier.  Per job a warp overlaps, behind its arithmetic:
// the claim of the job after next (atomic), the flag loads and then the cp.async prefetch (L2 path,
// 16-byte) of the next job's window into its staging buffer, and the release of the previous job's flag
// (its stores were issued a whole chunk earlier, so the release fence has nothing left to wait for).
// Results leave through a second staging buffer so that L2/HBM sees coalesced 16-byte stores.
//
// ONE WARP PER CTA: everything that steers the job loop then derives from blockIdx-uniform values,
// which lets ptxas keep the RK4 coefficients in uniform registers inside the loop -- with several warps
// per CTA the job number depends on threadIdx and it falls back to register operands (a DFMA with three
// register operands issues at 2/3 rate: the register file has two banks).
//
// Requirements (checked by l96_v2_eligible; otherwise the first design runs): even n and row strides,
// 16-byte aligned rows, no trajectory output.
#include <cstdlib>

#include "../../dataassimbench_b200/csrc/common.cuh"
#include "../../dataassimbench_b200/csrc/kernels.h"
#include "../../dataassimbench_b200/csrc/l96_dataflow.cuh"

namespace dab {

constexpr int V2_P = 16;              // points per thread
constexpr int V2_W = 32 * V2_P;       // points per window
constexpr int V2_STAGE = V2_W + 2 * (V2_W / 16);   // padded staging doubles per buffer (conflict-free LDS.128)
// The jobs of chunk c-1 a job depends on.  RAW: their output intersects my input window; WAR: their
// input window intersects the range I overwrite (same buffer, two chunks apart).  Both are covered by
// the jobs whose output range meets [p - 8 s, q + 8 s), s the larger of the two chunks' steps: <= 3 jobs
// unless the slab is much shorter than a window.  Every lane reads the SAME flag addresses (volatile
// loads, served by L2), so the outcome is warp-uniform by construction.  `v2_deps_load` only ISSUES the
// loads; the values are tested a step of arithmetic later.  The data the flags guard is read through
// cp.async.cg (L2 as well), issued after the branch on the flag values.
__device__ __forceinline__ V2Deps v2_deps_load(const L96V2Args& a, const V2Job& jb) {
  V2Deps d; d.v0 = d.v1 = d.v2 = a.epoch; d.many = false;
  if (jb.c == 0) return d;
  int j0, cnt;
  v2_deps_range(a, jb, j0, cnt);
  const int Jp = a.ch[jb.c - 1].J;
  d.many = cnt > 3;
  const unsigned* f0 = a.flags + a.ch[jb.c - 1].job_base + (long long)jb.m * Jp;
  d.v0 = v2_flag(f0, j0, 0, cnt, Jp); d.v1 = v2_flag(f0, j0, 1, cnt, Jp); d.v2 = v2_flag(f0, j0, 2, cnt, Jp);
  return d;
}

__device__ __forceinline__ bool v2_deps_test(const L96V2Args& a, const V2Deps& d) {
  return d.v0 == a.epoch && d.v1 == a.epoch && d.v2 == a.epoch && !d.many;
}

// blocking form (also the only form for > 3 dependencies)
__device__ __forceinline__ void v2_deps_wait(const L96V2Args& a, const V2Job& jb) {
  if (jb.c == 0) return;
  int j0, cnt;
  v2_deps_range(a, jb, j0, cnt);
  const int Jp = a.ch[jb.c - 1].J;
  const unsigned* f0 = a.flags + a.ch[jb.c - 1].job_base + (long long)jb.m * Jp;
  for (int i = 0; i < cnt; ++i)
    while (v2_flag(f0, j0, i, cnt, Jp) != a.epoch) { }
}

// release = gpu-scope fence (by every lane: each orders its own stores) + warp barrier + relaxed flag store
__device__ __forceinline__ void v2_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void v2_publish(const L96V2Args& a, int n, int lane) {
  __syncwarp();                          // every lane's fence happens before lane 0's flag store
  if (lane == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.flags + n), "r"(a.epoch) : "memory");
}

// cp.async (L2 path) of the job's 512-point window into the padded staging buffer: 256 16-byte pairs,
// lane l takes pairs l, l + 32, ...  Windows that lie inside the valid range (all but the ones at the
// ends of a row) need no clamping or wrapping: one address, eight immediate offsets.
__device__ __forceinline__ void v2_prefetch(const L96V2Args& a, const V2Job& jb, double* stage, int lane) {
  const double* row;
  int lo, hi;
  if (jb.c == 0) { row = a.in + (long long)jb.m * a.in_ld + a.in_base; lo = a.in_lo; hi = a.in_hi; }
  else {
    row = a.scratch[(jb.c - 1) & 1] + (long long)jb.m * a.scr_ld + a.scr_base;
    lo = a.ch[jb.c - 1].org; hi = a.ch[jb.c - 1].end;
  }
  if (a.periodic) { lo = 0; hi = a.n; }
  const unsigned dst0 = v2_smem(stage + 2 * lane + 2 * (lane >> 3));
  if (jb.g0 >= lo && jb.g0 + V2_W <= hi) {
    const double* src = row + jb.g0 + 2 * lane;
#pragma unroll
    for (int i = 0; i < V2_W / 64; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 72 * 8 * i), "l"(src + 64 * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < V2_W / 64; ++i) {
      int g = jb.g0 + 2 * lane + 64 * i;
      if (a.periodic) g = v2_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 72 * 8 * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

#ifdef V2_DEBUG_TIMING
__device__ long long g_v2_t[8];
__device__ long long g_v2_cnt[4];
#define V2_T(i) do { const long long t_ = clock64(); if (blockIdx.x == 300 && threadIdx.x == 0) g_v2_t[i] += t_ - tl_; tl_ = t_; } while (0)
#else
#define V2_T(i) do { } while (0)
#endif

#ifndef V2_MIN_CTAS
#define V2_MIN_CTAS 12
#endif
__global__ void __launch_bounds__(32, V2_MIN_CTAS)
l96_v2_kernel(const L96V2Args a) {
#ifdef V2_DEBUG_TIMING
  long long tl_ = clock64();
#endif
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x;
  double* st_in = smem;
  double* st_out = st_in + V2_STAGE;
  const int total = a.total;
  pdl_wait();
  pdl_launch_dependents();
  // claims: lane 0 adds to the dispenser, the value is broadcast when it is needed (a job later)
  auto claim_issue = [&]() -> unsigned long long {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(a.counter, 1ull) - a.ctr_base;
    return v;
  };
  auto claim_get = [&](unsigned long long v) -> int {
    v = __shfl_sync(0xffffffffu, v, 0);
    return v < (unsigned long long)total ? (int)v : total;
  };
  int n0 = claim_get(claim_issue());
  if (n0 >= total) return;
  V2Job cur = v2_decode(a, n0, 0);
  v2_deps_wait(a, cur);
  v2_prefetch(a, cur, st_in, lane);
  int n1 = claim_get(claim_issue());
  V2Job nxt = cur;
  if (n1 < total) nxt = v2_decode(a, n1, cur.c);
  // jobs whose stores are out but whose flags are not: published two at a time, so that the gpu-scope
  // fence (it stalls the warp for ~700 cycles whether or not stores are outstanding) is paid once per two
  // jobs; before this warp blocks on anything it publishes what it holds
  int pend0 = -1, pend1 = -1;
  auto flush = [&]() {
    if (pend0 >= 0) {
      __syncwarp();
      v2_fence();
      v2_publish(a, pend0, lane);
      if (pend1 >= 0) v2_publish(a, pend1, lane);
      pend0 = pend1 = -1;
    }
  };

  while (true) {
    V2_T(7);
    const bool have_next = n1 < total;
    // ---- behind this job's arithmetic: claim the job after next, look at the next job's producers
    unsigned long long claim = 0;
    V2Deps nd; nd.v0 = nd.v1 = nd.v2 = a.epoch; nd.many = false;
    if (have_next) { claim = claim_issue(); nd = v2_deps_load(a, nxt); }
    // ---- this job's window: staging -> registers (blocked: lane owns 16 consecutive points)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    double s[V2_P], xh[V2_P], xd[V2_P], acc[V2_P];
#pragma unroll
    for (int i = 0; i < V2_P; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(st_in + 18 * lane + i);
      s[i] = v.x; s[i + 1] = v.y;
    }
    __syncwarp();
    V2_T(0);
    // ---- the chunk: 17 FP64 instructions per point and step, neighbours by shuffle, no barrier
    bool fetched = false;
    const int ns = a.ch[cur.c].steps;
    for (int step = 0; step < ns; ++step) {
      if (step == 1 && have_next && v2_deps_test(a, nd)) {     // the flag loads had a step to come back
        v2_prefetch(a, nxt, st_in, lane);
        fetched = true;
      }
#pragma unroll
      for (int i = 0; i < V2_P; ++i) { xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const double Lx = __shfl_up_sync(0xffffffffu, s[V2_P - 2], 1), Ly = __shfl_up_sync(0xffffffffu, s[V2_P - 1], 1);
        const double R = __shfl_down_sync(0xffffffffu, s[0], 1);
        double kk[V2_P];
        kk[0] = fma(s[1] - Lx, Ly, -s[0]);
        kk[1] = fma(s[2] - Ly, s[0], -s[1]);
        kk[V2_P - 1] = fma(R - s[V2_P - 3], s[V2_P - 2], -s[V2_P - 1]);
#pragma unroll
        for (int i = 2; i < V2_P - 1; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
        for (int i = 0; i < V2_P; ++i) {
          if (st == 0) { acc[i] = fma(kk[i], a.h6, xd[i]); s[i] = fma(kk[i], a.h2, xh[i]); }
          else if (st == 1) { acc[i] = fma(kk[i], a.h3, acc[i]); s[i] = fma(kk[i], a.h2, xh[i]); }
          else if (st == 2) { acc[i] = fma(kk[i], a.h3, acc[i]); s[i] = fma(kk[i], a.dt, xd[i]); }
          else s[i] = fma(kk[i], a.h6, acc[i]);
        }
      }
    }
    V2_T(2);
    // ---- two jobs' stores are out (issued at least a whole chunk ago): publish them
    if (pend1 >= 0) flush();
    V2_T(4);
    // ---- registers -> staging -> coalesced 16-byte stores of the still-valid points [p, q)
#pragma unroll
    for (int i = 0; i < V2_P; i += 2) *reinterpret_cast<double2*>(st_out + 18 * lane + i) = make_double2(s[i], s[i + 1]);
    __syncwarp();
    {
      const bool last = cur.c == a.n_chunks - 1;
      double* row = last ? a.out + (long long)cur.m * a.out_ld
                         : a.scratch[cur.c & 1] + (long long)cur.m * a.scr_ld + a.scr_base;
      double* dst = row + cur.g0 + 2 * lane;
      const double* srcs = st_out + 2 * lane + 2 * (lane >> 3);
      const int e_lo = cur.p - cur.g0 - 2 * lane, e_hi = cur.q - cur.g0 - 2 * lane;   // pair 64 i is stored iff e_lo <= 64 i < e_hi
#pragma unroll
      for (int i = 0; i < V2_W / 64; ++i) {
        if (64 * i >= e_lo && 64 * i < e_hi)
          *reinterpret_cast<double2*>(dst + 64 * i) = *reinterpret_cast<const double2*>(srcs + 72 * i);
      }
    }
    __syncwarp();
    if (pend0 < 0) pend0 = cur.n; else pend1 = cur.n;
    V2_T(3);
#ifdef V2_DEBUG_TIMING
    if (blockIdx.x == 300 && threadIdx.x == 0) { g_v2_cnt[0] += 1; g_v2_cnt[1] += (have_next && !fetched) ? 1 : 0; g_v2_cnt[2] += nd.many ? 1 : 0; }
#endif
    if (!have_next) break;
    if (!fetched) {
      // its producers were not done a step into this job (or the chunk has a single step): publish what we
      // hold (a producer may be our own job), wait for them, fetch
      if (!v2_deps_test(a, v2_deps_load(a, nxt))) {
        flush();
        v2_deps_wait(a, nxt);
      }
      v2_prefetch(a, nxt, st_in, lane);
    }
    V2_T(5);
    const int n2 = claim_get(claim);
    cur = nxt;
    n1 = n2;
    if (n2 < total) nxt = v2_decode(a, n2, cur.c);
    V2_T(1);
  }
  flush();
}

// ------------------------------------------------------------------------------------------ host side
static bool l96_v2x_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj) {
  if (traj || n_steps < 1 || n_steps > 6 * V2_MAX_CHUNKS) return false;
  if ((n & 1) || (in_ld & 1) || (out_ld & 1) || (in_base & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (n < 4 || n + 12ll * n_steps >= (1ll << 30)) return false;
  return true;
}

// scratch row stride for windows of `n_steps` steps over slabs of n points (extended layout, even)
static long long l96_v2x_scratch_ld(long long n, int n_steps) { return (n + 12ll * n_steps + 2) & ~1ll; }

static size_t l96_v2x_flag_count(long long n, int k, int n_steps) {
  // upper bound: <= n_steps chunks, each over at most n + 12 n_steps points in jobs of >= 440 useful points
  // (fewer only when the whole row is shorter)
  const long long per = (n + 12ll * n_steps) / 256 + 2;
  return (size_t)(n_steps < V2_MAX_CHUNKS ? n_steps : V2_MAX_CHUNKS) * (size_t)k * (size_t)per;
}

cudaError_t l96_v2_launch(const double* in, double* out, double* scratch0, double* scratch1, long long scr_ld,
                          unsigned* flags, unsigned epoch, unsigned long long* counter, unsigned long long* ctr_base,
                          long long n, int k, int n_steps, int chunk_steps,
                          double dt, double F, long long in_ld, long long out_ld, bool periodic,
                          long long in_base, long long in_lo, long long in_hi, int sms, cudaStream_t st) {
  if (chunk_steps < 1) chunk_steps = 5;
  if (chunk_steps > 6) chunk_steps = 6;
  L96V2Args a{};
  a.in = in; a.out = out; a.scratch[0] = scratch0; a.scratch[1] = scratch1;
  a.in_ld = in_ld; a.out_ld = out_ld; a.scr_ld = scr_ld;
  a.in_base = in_base; a.in_lo = (int)in_lo; a.in_hi = (int)in_hi;
  if (!l96_dataflow_plan(a, n, k, n_steps, chunk_steps, V2_W, periodic)) return cudaErrorInvalidValue;
  const long long base = a.total;
  a.flags = flags; a.epoch = epoch;
  a.counter = counter; a.ctr_base = *ctr_base;
  a.dt = dt; a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  const size_t smem = 2 * (size_t)V2_STAGE * sizeof(double);
  static int per_sm = 0;                          // all CTAs must be co-resident: the flag waits rely on it
  if (per_sm == 0) {
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, l96_v2_kernel, 32, smem);
    if (e != cudaSuccess) return e;
    per_sm = occ < 1 ? 1 : occ;
    if (const char* ev = getenv("DAB_L96_V2_WARPS")) { const int v = atoi(ev); if (v >= 1 && v <= per_sm) per_sm = v; }
  }
  long long grid = (long long)sms * per_sm;
  if (grid > base) grid = base;
  if (grid < 1) grid = 1;
  // every warp claims until its first miss and never again, so this launch advances the dispenser by
  // exactly jobs + warps: the next launch starts there (no reset between launches, nothing to memset)
  *ctr_base += (unsigned long long)base + (unsigned long long)grid;
  return launch_chain(true, l96_v2_kernel, dim3((unsigned)grid), dim3(32), smem, st, a);
}

}  // namespace dab
