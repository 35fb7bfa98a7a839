// EXPERIMENT (not part of the product build since l96_forecast_v6.cu replaced it; tools/experiments/l96_v5_test.cu runs it):
// correct and bit-identical to the first design; C4 6.85 ms against 6.71 ms for the design that lets the groups run
// freely (the wavefront costs nothing by itself, but the select between the two sources of the left boundary and the
// second, predicated history store cost every warp issue slots: profiles/forecast_probes_r02.md section 6).
//
// K1, fifth design -- pipelined chains: one persistent CTA per SM, three 4-warp groups that sweep ONE chain of tiles
// together, each tile taking its left boundary from the tile before it through shared memory.
//
// Replaces the same reference functions as l96_forecast.cu (ETKF._step_forecast dabench/dacycler/_etkf.py:64-80
// -> Model.forecast -> Data.generate dabench/data/_data.py:86-211 -> integrate dabench/data/_utils.py:16-57 with
// Lorenz96.rhs dabench/data/_lorenz96.py:119-151), classical RK4 for all members at once.
//
// What round 2 measured (profiles/forecast_probes_r02.md, tools/l96_v4_test.cu):
//   * the inner loop of the first design (tile of 128 x 16 points in registers, edges through shared memory, one
//     barrier per RK stage) keeps the FP64 pipe ~85 % busy when three such tiles share an SM;
//   * a tile recomputes a 12 S-point halo (240 of 2048 points at S = 20), 8 S of it on the LEFT: the Lorenz-96
//     stencil is one-sided (i-2, i-1 | i+1);
//   * three co-resident CTAs do NOT share the pipe fairly: the warp schedulers favour the oldest CTA, which runs at
//     its single-CTA speed (14 us per tile) while the other two get the rest (19 and 25 us per tile).  With equal
//     static work per CTA the last one finishes alone at 38 % pipe utilisation (measured: CTA lifetimes
//     4.3 / 5.9 / 7.6 ms at C4); a dynamic dispenser (l96_forecast_v3.cu) hides that but needs independent jobs,
//     i.e. the full halo for every tile.
// This design removes the left halo AND the unfairness with one mechanism.  A CTA of 384 threads is three groups of
// 4 warps (own named barrier, own staging buffers); the CTA owns a contiguous range of the flattened (member,
// position) space, cut into tiles that advance by A = W - 4 S rounded down to whole threads (1968 of 2048 at
// S = 20: 3.9 % redundancy); group g takes tiles g, g + 3, g + 6, ...  Tile i records the two points just left of
// tile i+1's first point at EVERY RK stage of the window (4 S double2 of shared memory: the pair its thread
// `hist_t` publishes for its right-hand neighbour anyway), and thread 0 of tile i+1 reads that history where a
// fresh tile would read its own stale halo.  Tile i+1 may run step s once tile i has finished step s: one
// shared-memory progress word per tile, polled by one thread once per step just before the group's barrier.  The
// three groups therefore advance as a wavefront one step apart, a group that runs ahead waits for its predecessor
// instead of starving it, and all three finish within two steps of each other.  Only the first tile of a CTA's range
// and the first tile of every member pay the 8 S-point left halo.  Every value is computed by exactly the same
// instruction sequence as in the first design: results are bit-identical.
//
// Control flow is uniform by construction (every group runs the same number of rounds, taken from a kernel
// argument; a CTA with fewer tiles computes discarded dummy tiles in its last round), which is what lets ptxas keep
// the RK4 coefficients in uniform registers (see l96_forecast.cu).
//
// Requirements (l96_v5_eligible): even n and row strides, 16-byte aligned rows, no trajectory output.
#include <cstdlib>
#include <vector>

#include "../../dataassimbench_b200/csrc/common.cuh"
#include "../../dataassimbench_b200/csrc/kernels.h"

namespace dab {

struct L96V5Plan {
  int grid = 0, rounds = 0, hist_t = 0;
  std::vector<int> table;          // [grid][3 rounds][4]: {2 member + continuing, g0, p, q}; p == q: dummy tile
};

constexpr int V5_P = 16;                          // points per thread
constexpr int V5_TG = 128;                        // threads per group
constexpr int V5_G = 3;                           // groups per CTA
constexpr int V5_W = V5_P * V5_TG;                // points per tile
constexpr int V5_STAGE = V5_W + 2 * (V5_W / 16);  // padded staging doubles (16-byte pairs stay aligned, LDS.128 conflict-free)
constexpr int V5_MAX_STEPS = 32;                  // history entries: 4 per step
constexpr int V5_RING = 4;                        // history buffers: tile i+4 may overwrite tile i's (tile i+1 is done)

struct L96V5Args {
  const double* in;
  double* out;
  long long in_ld, out_ld;      // member strides (doubles)
  long long in_base;            // column of slab coordinate 0 in an input row
  int in_lo, in_hi;             // EXT: valid slab-coordinate range of an input row
  int n;                        // PERIODIC: ring size; EXT: slab length
  int periodic;
  const int4* jobs;             // [grid][3 rounds]: {2 * member + continuing, g0, p, q}; p == q: dummy tile
  int rounds;
  int n_steps;
  int hist_t;                   // the thread whose last two points are the next tile's left boundary
  double h2, h3, h6, dt, h2F, dtF;    // RK4 coefficients as uniform operands (see l96_forecast.cu)
};

__device__ __forceinline__ unsigned v5_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void v5_group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(V5_TG) : "memory"); }

__device__ __forceinline__ int v5_wrap(int g, int n) {      // ring coordinate of g (any distance from [0, n))
  if ((unsigned)g < (unsigned)n) return g;
  if (g < 0) { g += n; if (g >= 0) return g; }
  else { g -= n; if (g < n) return g; }
  g %= n;
  return g < 0 ? g + n : g;
}

// cp.async (L2 path) of a tile into a group's staging buffer: 1024 16-byte pairs, thread t takes pairs t, t + 128, ...
__device__ __forceinline__ void v5_prefetch(const L96V5Args& a, const int4 jb, double* stage, int t) {
  const double* row = a.in + (long long)(jb.x >> 1) * a.in_ld + a.in_base;
  int lo = a.in_lo, hi = a.in_hi;
  if (a.periodic) { lo = 0; hi = a.n; }
  const int g0 = jb.y;
  // pair index q = t + 128 i -> element e = 2 q = 2 t + 256 i; padded position e + 2 (e >> 4) = 2 t + 2 (t >> 3) + 288 i
  const unsigned dst0 = v5_smem(stage + 2 * t + 2 * (t >> 3));
  if (g0 >= lo && g0 + V5_W <= hi) {
    const double* src = row + g0 + 2 * t;
#pragma unroll
    for (int i = 0; i < V5_W / (2 * V5_TG); ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(src + 256 * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < V5_W / (2 * V5_TG); ++i) {
      int g = g0 + 2 * t + 256 * i;
      if (a.periodic) g = v5_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(V5_G * V5_TG, 1)
l96_v5_kernel(const L96V5Args a) {
  constexpr int P = V5_P, T = V5_TG;
  extern __shared__ __align__(16) double smem[];
  const int grp = threadIdx.x / T;                      // 0..2
  const int t = threadIdx.x - grp * T;                  // thread of the group
  double* gbase = smem + (size_t)grp * (6 * T + 2 * V5_STAGE);
  double2* e_last2 = reinterpret_cast<double2*>(gbase);  // [2][T] (s_{P-2}, s_{P-1}) of every thread
  double* e_first = gbase + 4 * T;                       // [2][T] s_0 of every thread
  double* st_in = e_first + 2 * T;                       // prefetched tile of the group's next job
  double* st_out = st_in + V5_STAGE;                     // results on their way out
  double2* hist = reinterpret_cast<double2*>(smem + (size_t)V5_G * (6 * T + 2 * V5_STAGE));   // [V5_RING][4 V5_MAX_STEPS]
  __shared__ int s_prog[V5_RING];                        // 64 * (tile + 1) + steps the tile has finished
  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;
  const int4* jobs = a.jobs + (long long)blockIdx.x * (V5_G * a.rounds);
  if (threadIdx.x < V5_RING) s_prog[threadIdx.x] = 0;
  // the job table is written once at plan time, not by the previous kernel: read it before the wait
  int4 cur = __ldg(jobs + grp);
  __syncthreads();
  pdl_wait();                 // the input is the previous kernel's output
  pdl_launch_dependents();
  v5_prefetch(a, cur, st_in, t);
  const int nq = 4 * a.n_steps;

  for (int r = 0; r < a.rounds; ++r) {
    const int tile = V5_G * r + grp;                     // index in the CTA's chain
    // the descriptor of the group's next tile (a dummy repeat of this one after the last round)
    const int4 nxt = __ldg(jobs + (r + 1 < a.rounds ? tile + V5_G : tile));
    // ---- this tile: staging -> registers (blocked: thread owns 16 consecutive points)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    v5_group_sync(grp);
    double s[P], xh[P], xd[P], acc[P];
#pragma unroll
    for (int i = 0; i < P; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(st_in + 18 * t + i);
      s[i] = v.x; s[i + 1] = v.y;
    }
#pragma unroll
    for (int i = 0; i < P; ++i) { xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
    int buf = 0;
    e_first[t] = s[0];
    e_last2[t] = make_double2(s[P - 2], s[P - 1]);
    const double2* hist_in = hist + ((tile + V5_RING - 1) % V5_RING) * (4 * V5_MAX_STEPS);
    double2* hist_out = hist + (tile % V5_RING) * (4 * V5_MAX_STEPS);
    volatile int* prog_in = s_prog + (tile + V5_RING - 1) % V5_RING;
    volatile int* prog_out = s_prog + tile % V5_RING;
    const bool cont = cur.x & 1;                          // a continuing tile: the left boundary is the previous tile's
    const bool from_hist = (t == 0) && cont;
    const bool to_hist = t == a.hist_t;
    if (to_hist) hist_out[0] = make_double2(s[P - 2], s[P - 1]);

    int q = 0;
    for (int step = 0; step < a.n_steps; ++step) {
      // tile i may run step s once tile i-1 has finished it (its history entries 4 s .. 4 s + 3 are written)
      if (from_hist) { const int need = 64 * tile + step + 1; while (*prog_in < need) { } }
#pragma unroll
      for (int st = 0; st < 4; ++st, ++q) {
        v5_group_sync(grp);
        if (step == 0 && st == 1 && r + 1 < a.rounds) v5_prefetch(a, nxt, st_in, t);   // every thread left st_in a barrier ago
        const double2 L = from_hist ? hist_in[q] : e_last2[buf * T + tl];
        const double R = e_first[buf * T + tr];
        buf ^= 1;
#define V5_CA (st == 2 ? a.dt : a.h2)
#define V5_CW ((st == 0 || st == 3) ? a.h6 : a.h3)
        double kk[P];
        // edge points first so the next stage's edges can be published early
        kk[0] = fma(s[1] - L.x, L.y, -s[0]);
        kk[1] = fma(s[2] - L.y, s[0], -s[1]);
        kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
        kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
        {
          double n0, n1, n2;
          if (st < 3) {
            n0 = fma(V5_CA, kk[0], st == 2 ? xd[0] : xh[0]);
            n2 = fma(V5_CA, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
            n1 = fma(V5_CA, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
          } else {
            n0 = fma(V5_CW, kk[0], acc[0]);
            n2 = fma(V5_CW, kk[P - 2], acc[P - 2]);
            n1 = fma(V5_CW, kk[P - 1], acc[P - 1]);
          }
          e_first[buf * T + t] = n0;
          e_last2[buf * T + t] = make_double2(n2, n1);
          if (to_hist && q + 1 < nq) hist_out[q + 1] = make_double2(n2, n1);
        }
#pragma unroll
        for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (st == 0) acc[i] = fma(V5_CW, kk[i], xd[i]);
          else if (st < 3) acc[i] = fma(V5_CW, kk[i], acc[i]);
          if (st < 3) s[i] = fma(V5_CA, kk[i], st == 2 ? xd[i] : xh[i]);
          else { s[i] = fma(V5_CW, kk[i], acc[i]); xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
        }
#undef V5_CA
#undef V5_CW
      }
      // the history of this step is in shared memory: let the next tile run it
      if (to_hist) { __threadfence_block(); *prog_out = 64 * (tile + 1) + step + 1; }
    }
    // ---- registers -> staging -> coalesced 16-byte stores of the still-valid points [p, q)
#pragma unroll
    for (int i = 0; i < P; i += 2) *reinterpret_cast<double2*>(st_out + 18 * t + i) = make_double2(s[i], s[i + 1]);
    v5_group_sync(grp);
    {
      double* dst = a.out + (long long)(cur.x >> 1) * a.out_ld + cur.y + 2 * t;
      const double* srcs = st_out + 2 * t + 2 * (t >> 3);
      const int e_lo = cur.z - cur.y - 2 * t, e_hi = cur.w - cur.y - 2 * t;   // pair 256 i is stored iff e_lo <= 256 i < e_hi
#pragma unroll
      for (int i = 0; i < V5_W / (2 * T); ++i) {
        if (256 * i >= e_lo && 256 * i < e_hi)
          *reinterpret_cast<double2*>(dst + 256 * i) = *reinterpret_cast<const double2*>(srcs + 288 * i);
      }
    }
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------ host side
int l96_v5_tile() { return V5_W; }

bool l96_v5_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj) {
  if (traj || n_steps < 1 || n_steps > V5_MAX_STEPS || 12 * n_steps + 16 > V5_W) return false;
  if ((n & 1) || (in_ld & 1) || (out_ld & 1) || (in_base & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (n < 4 || n + 12ll * n_steps >= (1ll << 30)) return false;
  return true;
}

static size_t v5_smem_bytes() {
  return ((size_t)V5_G * (6 * (size_t)V5_TG + 2 * (size_t)V5_STAGE) + 2 * (size_t)V5_RING * 4 * V5_MAX_STEPS) * sizeof(double);
}

// Cut k rows of n points into <= ctas chains of <= m tiles, m minimal and a multiple of 3 (one round = one tile per
// group).  A chain walks the flattened (member, position) space.  With A = W - 4 S rounded down to whole threads,
// a tile whose window starts at g0 hands its successor the window g0 + A (the successor's left boundary is thread
// hist_t's last two points, hist_t = A / P - 1).  The first tile of a chain and the first tile of every member it
// enters are fresh: first useful point `pos`, window from pos - 8 S, useful [pos, g0 + A); the others continue:
// useful [g0, g0 + A).  `v5_walk` emits the jobs for one value of m (or just counts the chains when `table` is
// null); a chain shorter than m is padded with dummy tiles (p == q: computed, never stored).
static int v5_walk(long long n, int k, int S, int m, std::vector<int>* table) {
  const int A = ((V5_W - 4 * S) / V5_P) * V5_P;
  int chains = 0, in_chain = m;          // in_chain == m: the next tile opens a chain
  auto pad = [&]() {
    if (table)
      while ((long long)(table->size() / 4) < (long long)chains * m) {
        table->push_back(0); table->push_back(0); table->push_back(0); table->push_back(0);
      }
  };
  for (int mem = 0; mem < k; ++mem) {
    long long pos = 0;
    bool fresh = true;
    while (pos < n) {
      if (in_chain == m) { pad(); ++chains; in_chain = 0; fresh = true; }
      const long long g0 = fresh ? pos - 8 * S : pos;
      const long long q = g0 + A < n ? g0 + A : n;
      if (table) {
        table->push_back(2 * mem + (fresh ? 0 : 1));
        table->push_back((int)g0);
        table->push_back((int)pos);
        table->push_back((int)q);
      }
      pos = g0 + A;
      fresh = false;
      ++in_chain;
    }
  }
  pad();
  return chains;
}

bool l96_v5_plan(long long n, int k, int n_steps, int ctas, L96V5Plan* out) {
  if (n_steps < 1 || n_steps > V5_MAX_STEPS || 12 * n_steps + 16 > V5_W || ctas < 1) return false;
  const int A = ((V5_W - 4 * n_steps) / V5_P) * V5_P;
  long long m = ceil_div_ll(ceil_div_ll(n * k, (long long)ctas * A), V5_G) * V5_G;
  if (m < V5_G) m = V5_G;
  while (m < (1 << 20) && v5_walk(n, k, n_steps, (int)m, nullptr) > ctas) m += V5_G;
  if (m >= (1 << 20)) return false;
  out->table.clear();
  out->grid = v5_walk(n, k, n_steps, (int)m, &out->table);
  out->rounds = (int)(m / V5_G);
  out->hist_t = A / V5_P - 1;
  return true;
}

cudaError_t l96_v5_launch(const double* in, double* out, const int* jobs_dev, int grid, int rounds, int hist_t,
                          long long n, int n_steps, double dt, double F, long long in_ld, long long out_ld,
                          bool periodic, long long in_base, long long in_lo, long long in_hi, cudaStream_t st) {
  L96V5Args a{};
  a.in = in; a.out = out;
  a.in_ld = in_ld; a.out_ld = out_ld; a.in_base = in_base;
  a.in_lo = (int)in_lo; a.in_hi = (int)in_hi;
  a.n = (int)n; a.periodic = periodic ? 1 : 0;
  a.jobs = reinterpret_cast<const int4*>(jobs_dev); a.rounds = rounds;
  a.n_steps = n_steps; a.hist_t = hist_t;
  a.dt = dt; a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(l96_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v5_smem_bytes());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_chain(true, l96_v5_kernel, dim3((unsigned)grid), dim3(V5_G * V5_TG), v5_smem_bytes(), st, a);
}

}  // namespace dab
