// EXPERIMENT (not part of the product build; tools/experiments/l96_v4_test.cu runs it): correct and bit-identical to the
// first design, but no faster -- three co-resident CTAs do not share the FP64 pipe fairly, so equal static chains end
// with one CTA running alone (profiles/forecast_probes_r02.md); l96_forecast_v5.cu is the design that came out of it.
//
// K1, fourth design -- chained tiles: the persistent CTA-tile forecast kernel without the left halo.
//
// Replaces the same reference functions as l96_forecast.cu (ETKF._step_forecast dabench/dacycler/_etkf.py:64-80
// -> Model.forecast -> Data.generate dabench/data/_data.py:86-211 -> integrate dabench/data/_utils.py:16-57 with
// Lorenz96.rhs dabench/data/_lorenz96.py:119-151), classical RK4 for all members at once.
//
// What round 2 measured about the earlier designs (profiles/forecast_probes_r02.md): the inner loop (tile of
// 128 x 16 points in registers, edges through shared memory, a barrier per RK stage, three CTAs per SM) keeps the
// FP64 pipe ~85 % busy, and any finer-grained meeting of warps or CTAs costs more than it saves.  What is lost at
// C3 is (i) the 12 S = 240-point halo a tile recomputes (11.7 % of 2048 points; 15.6 % after balancing 37 tiles per
// member), (ii) wave quantisation: 2368 tiles over 444 resident CTAs = 5.33 waves, paid as 6.
//
// The Lorenz-96 stencil is one-sided: (i-2, i-1 | i+1), so 8 S of the 12 S halo points are on the LEFT.  A CTA
// that sweeps a member from left to right can hand the left boundary from one tile to the next instead of
// recomputing it: tile j records the two points just left of tile j+1's first point at EVERY RK stage of the
// window (4 S double2 = 1.3 KB of shared memory at S = 20: exactly the pair its thread `hist_t` publishes for its
// right-hand neighbour anyway), and thread 0 of tile j+1 reads that history where a fresh tile reads its own stale
// halo.  A continuing tile therefore advances by A = W - 4 S rounded down to whole threads (1968 of 2048 at
// S = 20: 3.9 % redundancy), only the first tile of a chain pays the full 12 S halo, nothing is exchanged between
// CTAs, and every value is computed by exactly the same instruction sequence as in the first design (results are
// bit-identical).
//
// Chains: the launch is persistent (one CTA per resident slot: 3 per SM); the host cuts the k rows into at most
// `slots` chains of at most m tiles each, m minimal (l96_v4_plan; a chain may run from one member into the
// next, starting a fresh tile there), and uploads the job table once per shape.  At C3 that is 444 chains of
// 5 tiles (2172 tiles instead of 2368, and no partial wave).  The next tile is prefetched with cp.async into a
// second staging buffer while the current one computes; results leave as coalesced 16-byte stores.
//
// Requirements (l96_v4_eligible): even n and row strides, 16-byte aligned rows, no trajectory output.
#include <cstdlib>
#include <vector>

#include "../../dataassimbench_b200/csrc/common.cuh"
#include "../../dataassimbench_b200/csrc/kernels.h"

namespace dab {

struct L96V4Plan {
  int grid = 0, jobs_per_cta = 0, hist_t = 0;
  std::vector<int> table;          // [grid][jobs_per_cta][4]: {2 member + continuing (< 0: none), g0, p, q}
};

constexpr int V4_P = 16;                          // points per thread
constexpr int V4_T = 128;                         // threads per CTA
constexpr int V4_W = V4_P * V4_T;                 // points per tile
constexpr int V4_STAGE = V4_W + 2 * (V4_W / 16);  // padded staging doubles (16-byte pairs stay aligned, LDS.128 conflict-free)
constexpr int V4_MAX_STEPS = 32;                  // history entries: 4 per step

struct L96V4Args {
  const double* in;
  double* out;
  long long in_ld, out_ld;      // member strides (doubles)
  long long in_base;            // column of slab coordinate 0 in an input row
  int in_lo, in_hi;             // EXT: valid slab-coordinate range of an input row
  int n;                        // PERIODIC: ring size; EXT: slab length
  int periodic;
  const int4* jobs;             // [grid][jobs_per_cta]: {2 * member + continuing (< 0: none), g0, p, q}
  int jobs_per_cta;
  int n_steps;
  int hist_t;                   // the thread whose last two points are the next tile's left boundary
  int stagger_div;              // debug/tuning: CTAs of class blockIdx / stagger_div start class * stagger_ns late
  unsigned stagger_ns;
  double h2, h3, h6, dt, h2F, dtF;    // RK4 coefficients as uniform operands (see l96_forecast.cu)
};

__device__ __forceinline__ unsigned v4_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int v4_wrap(int g, int n) {      // ring coordinate of g (any distance from [0, n))
  if ((unsigned)g < (unsigned)n) return g;
  if (g < 0) { g += n; if (g >= 0) return g; }
  else { g -= n; if (g < n) return g; }
  g %= n;
  return g < 0 ? g + n : g;
}

#ifndef V4_DIRECT_IO
#define V4_DIRECT_IO 1
#endif
#if V4_DIRECT_IO
// cp.async (L2 path) of a job's tile into the staging buffer, thread-private: thread t fetches its own 16
// consecutive points (8 16-byte pairs) to the place it will read them from, so no barrier is needed around the
// staging buffer.  A warp-level copy touches 32 half-used sectors; the other halves follow in the next copy.
__device__ __forceinline__ void v4_prefetch(const L96V4Args& a, const int4 jb, double* stage, int t) {
  const double* row = a.in + (long long)(jb.x >> 1) * a.in_ld + a.in_base;
  int lo = a.in_lo, hi = a.in_hi;
  if (a.periodic) { lo = 0; hi = a.n; }
  const int g0 = jb.y;
  const unsigned dst0 = v4_smem(stage + 18 * t);
  if (g0 >= lo && g0 + V4_W <= hi) {
    const double* src = row + g0 + 16 * t;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16 * i), "l"(src + 2 * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      int g = g0 + 16 * t + 2 * i;
      if (a.periodic) g = v4_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16 * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
#else
// cp.async (L2 path) of a job's tile into a staging buffer: 1024 16-byte pairs, thread t takes pairs t, t + 128, ...
__device__ __forceinline__ void v4_prefetch(const L96V4Args& a, const int4 jb, double* stage, int t) {
  const double* row = a.in + (long long)(jb.x >> 1) * a.in_ld + a.in_base;
  int lo = a.in_lo, hi = a.in_hi;
  if (a.periodic) { lo = 0; hi = a.n; }
  const int g0 = jb.y;
  // pair index q = t + 128 i -> element e = 2 q = 2 t + 256 i; padded position e + 2 (e >> 4) = 2 t + 2 (t >> 3) + 288 i
  const unsigned dst0 = v4_smem(stage + 2 * t + 2 * (t >> 3));
  if (g0 >= lo && g0 + V4_W <= hi) {
    const double* src = row + g0 + 2 * t;
#pragma unroll
    for (int i = 0; i < V4_W / (2 * V4_T); ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(src + 256 * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < V4_W / (2 * V4_T); ++i) {
      int g = g0 + 2 * t + 256 * i;
      if (a.periodic) g = v4_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

#endif

// -DV4_TIMING (tools/l96_v4_test.cu): thread 0 of every CTA accumulates the cycles it spends between the sections of a
// job {wait for the tile + registers, stage loop, stores} into g_v4_timing[8 blockIdx ..],
// with the CTA's start / end (globaltimer ns) and its lifetime in cycles
#ifdef V4_TIMING
__device__ long long* g_v4_timing = nullptr;
#define V4_CLK(var) const long long var = clock64()
#else
#define V4_CLK(var) do { } while (0)
#endif

__global__ void __launch_bounds__(V4_T, 3)
l96_v4_kernel(const L96V4Args a) {
#ifdef V4_TIMING
  long long tm_load = 0, tm_loop = 0, tm_store = 0, tm_jobs = 0, tm_gap = 0, tm_prev = 0, tm_g0, tm_c0 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm_g0));
#endif
  constexpr int P = V4_P, T = V4_T;
  extern __shared__ __align__(16) double smem[];
  double2* e_last2 = reinterpret_cast<double2*>(smem);  // [2][T] (s_{P-2}, s_{P-1}) of every thread
  double* e_first = smem + 4 * T;                       // [2][T] s_0 of every thread
  double* st_in = e_first + 2 * T;                      // prefetched tile of the next job
#if V4_DIRECT_IO
  double2* hist = reinterpret_cast<double2*>(st_in + V4_STAGE);    // [2][4 V4_MAX_STEPS] left-boundary history
#else
  double* st_out = st_in + V4_STAGE;                    // results on their way out
  double2* hist = reinterpret_cast<double2*>(st_out + V4_STAGE);   // [2][4 V4_MAX_STEPS] left-boundary history
#endif
  __shared__ int4 s_job[2];
  const int t = threadIdx.x;
  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;
  const int4* jobs = a.jobs + (long long)blockIdx.x * a.jobs_per_cta;
  // the job table is written once at plan time, not by the previous kernel: read it before the wait.  The job
  // descriptors reach the threads through shared memory (a load at a uniform address): the loop stays uniform to
  // the compiler, which keeps the RK4 coefficients in uniform registers (see l96_forecast.cu).
  if (t == 0) {
    s_job[0] = jobs[0];
    s_job[1] = a.jobs_per_cta > 1 ? jobs[1] : make_int4(-1, 0, 0, 0);
  }
  __syncthreads();
  int4 cur = s_job[0];
  pdl_wait();                 // the input is the previous kernel's output
  pdl_launch_dependents();
  if (cur.x < 0) return;
  if (a.stagger_ns != 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long wait = (unsigned long long)(blockIdx.x / a.stagger_div) * a.stagger_ns;
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < wait);
  }
  v4_prefetch(a, cur, st_in, t);
  const int nq = 4 * a.n_steps;

  for (int j = 0;; ++j) {
    // ---- this job's tile: staging -> registers (blocked: thread owns 16 consecutive points)
    V4_CLK(c_top);
#ifdef V4_TIMING
    if (tm_prev != 0) tm_gap += c_top - tm_prev;
#endif
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#if !V4_DIRECT_IO
    __syncthreads();
#endif
    double s[P], xh[P], xd[P], acc[P];
#pragma unroll
    for (int i = 0; i < P; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(st_in + 18 * t + i);
      s[i] = v.x; s[i + 1] = v.y;
    }
    // thread 0 fetches the descriptor after next now and parks it in shared memory at the end of this job
    int4 after = make_int4(-1, 0, 0, 0);
    if (t == 0 && j + 2 < a.jobs_per_cta) after = __ldg(jobs + j + 2);
#pragma unroll
    for (int i = 0; i < P; ++i) { xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
    int buf = 0;
    e_first[t] = s[0];
    e_last2[t] = make_double2(s[P - 2], s[P - 1]);
    const double2* hist_in = hist + (j & 1) * (4 * V4_MAX_STEPS);
    double2* hist_out = hist + ((j + 1) & 1) * (4 * V4_MAX_STEPS);
    const bool from_hist = (t == 0) && (cur.x & 1);      // a continuing tile: the left boundary is the previous tile's
    const bool to_hist = t == a.hist_t;
    if (to_hist) hist_out[0] = make_double2(s[P - 2], s[P - 1]);
    int4 nxt = make_int4(-1, 0, 0, 0);
    V4_CLK(c_loop);

    int q = 0;
    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
      for (int st = 0; st < 4; ++st, ++q) {
        __syncthreads();
        if (step == 0 && st == 0) {
          // the next job's descriptor (parked by thread 0 before this barrier) and the prefetch of its tile: this
          // thread has taken its points out of the staging buffer
          nxt = s_job[(j + 1) & 1];
#if V4_DIRECT_IO
          if (nxt.x >= 0) v4_prefetch(a, nxt, st_in, t);
#endif
        }
#if !V4_DIRECT_IO
        if (step == 0 && st == 1 && nxt.x >= 0) v4_prefetch(a, nxt, st_in, t);   // every thread left st_in a barrier ago
#endif
        const double2 L = from_hist ? hist_in[q] : e_last2[buf * T + tl];
        const double R = e_first[buf * T + tr];
        buf ^= 1;
#define V4_CA (st == 2 ? a.dt : a.h2)
#define V4_CW ((st == 0 || st == 3) ? a.h6 : a.h3)
        double kk[P];
        // edge points first so the next stage's edges can be published early
        kk[0] = fma(s[1] - L.x, L.y, -s[0]);
        kk[1] = fma(s[2] - L.y, s[0], -s[1]);
        kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
        kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
        {
          double n0, n1, n2;
          if (st < 3) {
            n0 = fma(V4_CA, kk[0], st == 2 ? xd[0] : xh[0]);
            n2 = fma(V4_CA, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
            n1 = fma(V4_CA, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
          } else {
            n0 = fma(V4_CW, kk[0], acc[0]);
            n2 = fma(V4_CW, kk[P - 2], acc[P - 2]);
            n1 = fma(V4_CW, kk[P - 1], acc[P - 1]);
          }
          e_first[buf * T + t] = n0;
          e_last2[buf * T + t] = make_double2(n2, n1);
          if (to_hist && q + 1 < nq) hist_out[q + 1] = make_double2(n2, n1);
        }
#pragma unroll
        for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (st == 0) acc[i] = fma(V4_CW, kk[i], xd[i]);
          else if (st < 3) acc[i] = fma(V4_CW, kk[i], acc[i]);
          if (st < 3) s[i] = fma(V4_CA, kk[i], st == 2 ? xd[i] : xh[i]);
          else { s[i] = fma(V4_CW, kk[i], acc[i]); xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
        }
#undef V4_CA
#undef V4_CW
      }
    }
    // ---- the still-valid points [p, q) leave
    V4_CLK(c_store);
#if V4_DIRECT_IO
    {
      // straight from the registers: 8 16-byte stores per thread (a warp-level store fills half of 32 sectors, the
      // next one the other half; L2 merges them)
      const int g = cur.y + 16 * t;
      double* dst = a.out + (long long)(cur.x >> 1) * a.out_ld + g;
#pragma unroll
      for (int i = 0; i < P; i += 2)
        if (g + i >= cur.z && g + i < cur.w) *reinterpret_cast<double2*>(dst + i) = make_double2(s[i], s[i + 1]);
    }
#else
#pragma unroll
    for (int i = 0; i < P; i += 2) *reinterpret_cast<double2*>(st_out + 18 * t + i) = make_double2(s[i], s[i + 1]);
    __syncthreads();
    {
      double* dst = a.out + (long long)(cur.x >> 1) * a.out_ld + cur.y + 2 * t;
      const double* srcs = st_out + 2 * t + 2 * (t >> 3);
      const int e_lo = cur.z - cur.y - 2 * t, e_hi = cur.w - cur.y - 2 * t;   // pair 256 i is stored iff e_lo <= 256 i < e_hi
#pragma unroll
      for (int i = 0; i < V4_W / (2 * T); ++i) {
        if (256 * i >= e_lo && 256 * i < e_hi)
          *reinterpret_cast<double2*>(dst + 256 * i) = *reinterpret_cast<const double2*>(srcs + 288 * i);
      }
    }
#endif
    // the descriptor after next: its slot was last read at the first barrier of THIS job, and is read next after
    // the first barrier of the next one
    if (t == 0) s_job[j & 1] = after;
#ifdef V4_TIMING
    const long long c_end = clock64();
    tm_load += c_loop - c_top; tm_loop += c_store - c_loop; tm_store += c_end - c_store; ++tm_jobs; tm_prev = c_end;
    if (nxt.x < 0 && t == 0 && g_v4_timing) {
      long long* o = g_v4_timing + 8 * blockIdx.x;
      long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      o[0] = tm_load; o[1] = tm_loop; o[2] = tm_store; o[3] = tm_jobs; o[4] = tm_g0; o[5] = g1; o[6] = c_end - tm_c0; o[7] = tm_gap;
    }
#endif
    if (nxt.x < 0) break;
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------ host side
int l96_v4_tile() { return V4_W; }

bool l96_v4_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj) {
  if (traj || n_steps < 1 || n_steps > V4_MAX_STEPS || 12 * n_steps + 16 > V4_W) return false;
  if ((n & 1) || (in_ld & 1) || (out_ld & 1) || (in_base & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (n < 4 || n + 12ll * n_steps >= (1ll << 30)) return false;
  return true;
}

static size_t v4_smem_bytes() {
  return (6 * (size_t)V4_T + (V4_DIRECT_IO ? 1 : 2) * (size_t)V4_STAGE + 2 * 2 * 4 * (size_t)V4_MAX_STEPS) * sizeof(double);
}

int l96_v4_slots(int sms) {
  static int per_sm = 0;
  if (per_sm == 0) {
    if (cudaFuncSetAttribute(l96_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v4_smem_bytes()) != cudaSuccess)
      return sms;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, l96_v4_kernel, V4_T, v4_smem_bytes()) != cudaSuccess) return sms;
    per_sm = occ < 1 ? 1 : occ;
  }
  return sms * per_sm;
}

// Cut k rows of n points into <= slots chains of <= m tiles, m minimal.  A chain walks the flattened
// (member, position) space.  With A = W - 4 S rounded down to whole threads, a tile whose window starts at g0
// hands its successor the window g0 + A (the successor's left boundary is thread hist_t's last two points,
// hist_t = A / P - 1).  The first tile of a chain and the first tile of every member it enters are fresh: first
// useful point `pos`, window from pos - 8 S, useful [pos, g0 + A); the others continue: useful [g0, g0 + A).
// `v4_walk` emits the jobs for one value of m (or just counts the chains when `table` is null).
// chain c holds `len(c)` tiles; the table is padded to `stride` jobs per chain
template <typename LenFn>
static int v4_walk(long long n, int k, int S, LenFn len, int stride, std::vector<int>* table) {
  const int A = ((V4_W - 4 * S) / V4_P) * V4_P;
  int chains = 0, in_chain = 0, cur_len = 0;          // in_chain == cur_len: the next tile opens a chain
  static const bool nochain = getenv("DAB_L96_V4_NOCHAIN") != nullptr;   // experiment: every tile fresh
  auto pad = [&]() {
    if (table)
      while ((long long)(table->size() / 4) < (long long)chains * stride) {
        table->push_back(-1); table->push_back(0); table->push_back(0); table->push_back(0);
      }
  };
  for (int mem = 0; mem < k; ++mem) {
    long long pos = 0;
    bool fresh = true;
    while (pos < n) {
      if (in_chain == cur_len) { pad(); cur_len = len(chains); ++chains; in_chain = 0; fresh = true; }
      const long long g0 = fresh ? pos - 8 * S : pos;
      const long long q = g0 + A < n ? g0 + A : n;
      if (table) {
        table->push_back(2 * mem + (fresh ? 0 : 1));
        table->push_back((int)g0);
        table->push_back((int)pos);
        table->push_back((int)q);
      }
      pos = g0 + A;
      fresh = nochain;
      ++in_chain;
    }
  }
  pad();
  return chains;
}

bool l96_v4_plan(long long n, int k, int n_steps, int slots, L96V4Plan* out) {
  if (n_steps < 1 || n_steps > V4_MAX_STEPS || 12 * n_steps + 16 > V4_W || slots < 1) return false;
  const int A = ((V4_W - 4 * n_steps) / V4_P) * V4_P;
  out->table.clear();
  out->hist_t = A / V4_P - 1;
  // experiment: DAB_V4_CLASS_M="a,b,c" gives the chains of the three launch-order classes (blockIdx / 148: the
  // oldest CTA of an SM is favoured by the warp schedulers) a, b and c tiles
  if (const char* ev = getenv("DAB_V4_CLASS_M")) {
    int M[3] = {0, 0, 0};
    if (sscanf(ev, "%d,%d,%d", &M[0], &M[1], &M[2]) == 3 && M[0] > 0 && M[1] > 0 && M[2] > 0) {
      const int per = slots / 3;
      const int stride = M[0] > M[1] ? (M[0] > M[2] ? M[0] : M[2]) : (M[1] > M[2] ? M[1] : M[2]);
      auto len = [&](int c) { return M[c / per < 3 ? c / per : 2]; };
      out->grid = v4_walk(n, k, n_steps, len, stride, &out->table);
      out->jobs_per_cta = stride;
      return out->grid <= slots;
    }
  }
  long long m = ceil_div_ll(n * k, (long long)slots * A);
  if (m < 1) m = 1;
  auto fixed = [&](int) { return (int)m; };
  while (m < (1 << 20) && v4_walk(n, k, n_steps, fixed, (int)m, nullptr) > slots) ++m;
  if (m >= (1 << 20)) return false;
  out->grid = v4_walk(n, k, n_steps, fixed, (int)m, &out->table);
  out->jobs_per_cta = (int)m;
  return true;
}

cudaError_t l96_v4_launch(const double* in, double* out, const int* jobs_dev, int grid, int jobs_per_cta, int hist_t,
                          long long n, int k, int n_steps, double dt, double F, long long in_ld, long long out_ld,
                          bool periodic, long long in_base, long long in_lo, long long in_hi, cudaStream_t st) {
  (void)k;
  L96V4Args a{};
  a.in = in; a.out = out;
  a.in_ld = in_ld; a.out_ld = out_ld; a.in_base = in_base;
  a.in_lo = (int)in_lo; a.in_hi = (int)in_hi;
  a.n = (int)n; a.periodic = periodic ? 1 : 0;
  a.jobs = reinterpret_cast<const int4*>(jobs_dev); a.jobs_per_cta = jobs_per_cta;
  a.n_steps = n_steps; a.hist_t = hist_t;
  a.stagger_div = 148; a.stagger_ns = 0;
  if (const char* ev = getenv("DAB_L96_STAGGER_NS")) a.stagger_ns = (unsigned)atoi(ev);
  a.dt = dt; a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(l96_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v4_smem_bytes());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_chain(true, l96_v4_kernel, dim3((unsigned)grid), dim3(V4_T), v4_smem_bytes(), st, a);
}

}  // namespace dab
