// K1, chained tiles handed out inside the SM: one persistent 384-thread CTA per SM, three 4-warp groups, each
// sweeping its own chain of tiles; a tile takes its left boundary from the tile before it instead of recomputing a
// halo; chains of decreasing length come from a shared-memory dispenser.
//
// Replaces the same reference functions as l96_forecast.cu (ETKF._step_forecast dabench/dacycler/_etkf.py:64-80
// -> Model.forecast -> Data.generate dabench/data/_data.py:86-211 -> integrate dabench/data/_utils.py:16-57 with
// Lorenz96.rhs dabench/data/_lorenz96.py:119-151), classical RK4 for all members at once.
//
// What round 2 measured (profiles/forecast_probes_r02.md):
//   * the inner loop of the first design (tile of 128 x 16 points in registers, edges through shared memory, one
//     barrier per RK stage) keeps the FP64 pipe 80-85 % busy when three such tiles share an SM;
//   * a tile recomputes a 12 S-point halo (240 of 2048 points at S = 20), 8 S of it on the LEFT: the Lorenz-96
//     stencil is one-sided (i-2, i-1 | i+1);
//   * three tiles that share an SM do NOT share its pipe fairly: the warp schedulers favour the oldest warps, which
//     run at the speed they have alone (14 us per tile) while the others get the rest (19 and 25 us).  Equal static
//     shares therefore end with one tile running alone at 38 % pipe utilisation, and forcing equal progress (a
//     wavefront of dependent tiles, tools/experiments/l96_forecast_v5.cu) is no faster than letting them run freely.
// So: the CTA owns a contiguous range of the flattened (member, position) space, cut by the host into CHAINS -- a
// fresh tile (full 12 S halo) followed by continuing tiles that advance by A = W - 4 S rounded down to whole
// threads (1968 of 2048 at S = 20: 3.9 % redundancy instead of 11.7 %).  Tile j records the two points just left of
// tile j+1's first point at EVERY RK stage of the window (4 S double2 in the group's shared memory: the pair one of
// its threads publishes for its right-hand neighbour anyway), and thread 0 of tile j+1 reads that history where a
// fresh tile reads its own stale halo.  Chain lengths follow guided self-scheduling -- tiles left / 6, so long
// chains first and single tiles at the end -- and a group that finishes a chain takes the next one from a counter in
// shared memory (one atomicAdd per chain, no global traffic): the groups never wait for each other, the favoured
// one simply sweeps more tiles, all are busy until the last tiles, and only one tile in ~30 pays the left halo.
// The next tile is prefetched with cp.async into a second staging buffer while the current one computes; results
// leave as coalesced 16-byte stores.  Per value the instruction sequence is the one of the first design: results are
// bit-identical.
//
// ptxas keeps the RK4 coefficients in uniform registers (full-rate DFMA: see l96_forecast.cu) only where it can
// prove the warp converged; the trip count of the tile loop differs from group to group, so every pass starts with
// the group's barrier, which restores that proof.
//
// Requirements (l96_v6_eligible): even n and row strides, 16-byte aligned rows, no trajectory output.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace dab {

// Two builds: 3 groups x 128 threads (tiles of 2048 points: the large shapes) and 6 groups x 64 threads (tiles of 1024
// points: slabs of a grid-sharded run, where an SM gets only a few thousand points and finer tiles balance better).
constexpr int V6_P = 16;                          // points per thread
constexpr int V6_MAX_STEPS = 32;                  // history entries: 4 per step
__host__ __device__ constexpr int v6_tile(int TG) { return V6_P * TG; }                                  // points per tile
__host__ __device__ constexpr int v6_stage(int TG) { return v6_tile(TG) + 2 * (v6_tile(TG) / 16); }     // padded staging doubles (16-byte pairs stay aligned, LDS.128 conflict-free)
__host__ __device__ constexpr int v6_group_doubles(int TG) { return 8 * TG + 2 * TG + 2 * v6_stage(TG) + 2 * 4 * TG * 2; }   // shared memory of a group

struct L96V6Args {
  const double* in;
  double* out;
  long long in_ld, out_ld;      // member strides (doubles)
  long long in_base;            // column of slab coordinate 0 in an input row
  int in_lo, in_hi;             // EXT: valid slab-coordinate range of an input row
  int n;                        // PERIODIC: ring size; EXT: slab length
  int periodic;
  const int4* tiles;            // [grid][tile_stride]: {2 * member + continuing, g0, p, q}, in sweep order
  const int* chains;            // [grid][chain_stride]: n_chains, then the index of every chain's first tile, then n_tiles
  int tile_stride, chain_stride;
  int n_steps;
  int hist_t;                   // the thread whose last two points are the next tile's left boundary
  double h2, h3, h6, dt, h2F, dtF;    // RK4 coefficients as uniform operands (see l96_forecast.cu)
  ObsEmit emit;                 // optional compact copy of the observed columns (kernels.h)
};

__device__ __forceinline__ unsigned v6_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int TG>
__device__ __forceinline__ void v6_group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TG) : "memory"); }

__device__ __forceinline__ int v6_wrap(int g, int n) {      // ring coordinate of g (any distance from [0, n))
  if ((unsigned)g < (unsigned)n) return g;
  if (g < 0) { g += n; if (g >= 0) return g; }
  else { g -= n; if (g < n) return g; }
  g %= n;
  return g < 0 ? g + n : g;
}

// cp.async (L2 path) of a tile into a group's staging buffer: 8 TG 16-byte pairs, thread t takes pairs t, t + TG, ...
template <int TG>
__device__ __forceinline__ void v6_prefetch(const L96V6Args& a, const int4 jb, double* stage, int t) {
  const double* row = a.in + (long long)(jb.x >> 1) * a.in_ld + a.in_base;
  int lo = a.in_lo, hi = a.in_hi;
  if (a.periodic) { lo = 0; hi = a.n; }
  const int g0 = jb.y;
  // pair index q = t + TG i -> element e = 2 q = 2 t + 2 TG i; padded position e + 2 (e >> 4) = 2 t + 2 (t >> 3) + (2 TG + TG / 4) i
  constexpr int DSTEP = (2 * TG + TG / 4) * 8;       // bytes
  const unsigned dst0 = v6_smem(stage + 2 * t + 2 * (t >> 3));
  if (g0 >= lo && g0 + v6_tile(TG) <= hi) {
    const double* src = row + g0 + 2 * t;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + DSTEP * i), "l"(src + 2 * TG * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      int g = g0 + 2 * t + 2 * TG * i;
      if (a.periodic) g = v6_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + DSTEP * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// candidate rows of a tile's stored range [p, q) = [jb.z, jb.w): at most the tile's width + 126
__device__ __forceinline__ void v6_rows(const ObsEmit& em, const int4 jb, int* lo, int* hi) {
  int l = 0, h = 0;
  if (jb.w > jb.z) {
    l = __ldg(em.row_start + (jb.z >> 6));
    h = __ldg(em.row_start + ((jb.w - 1) >> 6) + 1);
    if (h > em.n_rows) h = em.n_rows;
  }
  *lo = l; *hi = h;
}

template <int TG, int G>
__global__ void __launch_bounds__(G * TG, 1)
l96_v6_kernel(const L96V6Args a) {
  constexpr int P = V6_P, T = TG, V6_G = G, V6_STAGE = v6_stage(TG), V6_GROUP_DOUBLES = v6_group_doubles(TG);
  extern __shared__ __align__(16) double smem[];
  const int grp = threadIdx.x / T;                      // 0..2
  const int t = threadIdx.x - grp * T;                  // thread of the group
  // Left-neighbour pairs (s_{P-2}, s_{P-1}) live in FOUR buffers, one per RK stage of a step, T double2 = 2048 bytes
  // apart, and the boundary history uses the same geometry -- hist[buffer][stage][step], stages 2048 bytes apart --
  // so that "thread t reads its left neighbour's pair" and "thread 0 of a continuing tile reads the previous tile's
  // history" (and the matching stores) are the SAME instructions with the same immediate offsets: a per-thread base
  // pointer that advances by 0 or 16 bytes per step.  (A select between the two sources and a predicated second
  // store cost 4 % of the kernel: every warp pays the issue slots.)
  constexpr int STG = T * 16;                            // bytes between the stages of a step, in both layouts
  double* gbase = smem + (size_t)grp * V6_GROUP_DOUBLES;
  double2* e4 = reinterpret_cast<double2*>(gbase);       // [4][T] pairs of every thread, by stage
  double* e_first = gbase + 8 * T;                       // [2][T] s_0 of every thread
  double* st_in = e_first + 2 * T;                       // prefetched tile of the group's next job
  double* st_out = st_in + V6_STAGE;                     // results on their way out
  char* hist = reinterpret_cast<char*>(st_out + V6_STAGE);   // [2][4 stages][T slots] double2 (steps 0 .. S used)
  __shared__ int s_next_chain;                           // the dispenser
  __shared__ int s_nxt[V6_G][2];                         // per group: index of its next tile (-1: none), end of that tile's chain
  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;
  const int4* tiles = a.tiles + (long long)blockIdx.x * a.tile_stride;
  const int* chains = a.chains + (long long)blockIdx.x * a.chain_stride;
  const int n_chains = __ldg(chains);
  // the tables are written once at plan time, not by the previous kernel: read them before the wait
  if (threadIdx.x == 0) s_next_chain = V6_G;             // chains 0 .. G-1 go to the groups directly
  int ti = grp < n_chains ? __ldg(chains + 1 + grp) : -1;           // current tile
  int te = grp < n_chains ? __ldg(chains + 2 + grp) : -1;           // end of the current chain
  __syncthreads();
  pdl_wait();                 // the input is the previous kernel's output
  pdl_launch_dependents();
  if (ti < 0) return;
  int4 cur = __ldg(tiles + ti);
  v6_prefetch<TG>(a, cur, st_in, t);
  // compact observation copy (ObsEmit): the candidate rows of a tile -- those of the 64-column blocks its stored range
  // [p, q) meets -- are known one tile ahead (r_lo, r_hi: fetched next to the tile's descriptor), and their columns are
  // copied into shared memory while the tile computes; thread t copies and later reads rows r_lo + t + T i only, so
  // its own cp.async.wait_group is all the ordering the copy needs
  const bool emit = a.emit.yb != nullptr;
  int* s_cols = reinterpret_cast<int*>(smem + (size_t)V6_G * V6_GROUP_DOUBLES) + grp * (v6_tile(TG) + 128);
  int r_lo = 0, r_hi = 0, nr_lo = 0, nr_hi = 0;
  if (emit) v6_rows(a.emit, cur, &r_lo, &r_hi);

  for (int n_local = 0;; ++n_local) {
    // a barrier at the top of every pass tells ptxas the warps are converged in this loop, whose trip count differs
    // from group to group (without it the RK4 coefficients leave the uniform registers)
    v6_group_sync<TG>(grp);
    // ---- the next tile of this group: the chain's next one, or the first tile of the next chain from the dispenser
    if (t == 0) {
      int nx = ti + 1;
      if (nx >= te) {
        const int c = atomicAdd(&s_next_chain, 1);
        nx = c < n_chains ? __ldg(chains + 1 + c) : -1;
        s_nxt[grp][0] = nx;
        s_nxt[grp][1] = c < n_chains ? __ldg(chains + 2 + c) : -1;
      } else {
        s_nxt[grp][0] = nx;
        s_nxt[grp][1] = te;
      }
    }
    // ---- this tile: staging -> registers (blocked: thread owns 16 consecutive points)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    v6_group_sync<TG>(grp);
    double s[P], xh[P], xd[P], acc[P];
#pragma unroll
    for (int i = 0; i < P; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(st_in + 18 * t + i);
      s[i] = v.x; s[i + 1] = v.y;
    }
    if (emit) {
      for (int r = r_lo + t; r < r_hi; r += T)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(v6_smem(s_cols + (r - r_lo))), "l"(a.emit.loc + r) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int ti_next = s_nxt[grp][0], te_next = s_nxt[grp][1];
    int4 nxt = cur;
    if (ti_next >= 0) nxt = __ldg(tiles + ti_next);
#pragma unroll
    for (int i = 0; i < P; ++i) { xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
    int buf = 0;
    e_first[t] = s[0];
    char* hist_in = hist + (n_local & 1) * (4 * STG);
    char* hist_out = hist + ((n_local + 1) & 1) * (4 * STG);
#ifdef V6_NOHIST   /* timing experiment only: wrong results for continuing tiles */
    const bool from_hist = false, to_hist = false, after_hist = false;
#else
    const bool from_hist = (t == 0) && (cur.x & 1);      // a continuing tile: the left boundary is the previous tile's
    const bool to_hist = t == a.hist_t;                  // this thread's pair is the next tile's left boundary ...
    const bool after_hist = t == a.hist_t + 1;           // ... and this thread's: it reads it from the history too
#endif
    // where this thread reads its left pair of stage st of step s: Lp + st * STG (+ 16 bytes per step from the history)
    const char* Lp = from_hist ? hist_in : (after_hist ? hist_out : reinterpret_cast<const char*>(e4 + tl));
    const int Linc = (from_hist || after_hist) ? 16 : 0;
    // where it publishes its own pair for stage st: Qp + st * STG
    char* Qp = to_hist ? hist_out : reinterpret_cast<char*>(e4 + t);
    const int Qinc = to_hist ? 16 : 0;
    *reinterpret_cast<double2*>(Qp) = make_double2(s[P - 2], s[P - 1]);

    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        v6_group_sync<TG>(grp);
        if (step == 0 && st == 1 && ti_next >= 0) {
          v6_prefetch<TG>(a, nxt, st_in, t);   // every thread left st_in a barrier ago
          if (emit) v6_rows(a.emit, nxt, &nr_lo, &nr_hi);
        }
        const double2 L = *reinterpret_cast<const double2*>(Lp + st * STG);
        const double R = e_first[buf * T + tr];
        buf ^= 1;
#define V6_CA (st == 2 ? a.dt : a.h2)
#define V6_CW ((st == 0 || st == 3) ? a.h6 : a.h3)
        double kk[P];
        // edge points first so the next stage's edges can be published early
        kk[0] = fma(s[1] - L.x, L.y, -s[0]);
        kk[1] = fma(s[2] - L.y, s[0], -s[1]);
        kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
        kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
        {
          double n0, n1, n2;
          if (st < 3) {
            n0 = fma(V6_CA, kk[0], st == 2 ? xd[0] : xh[0]);
            n2 = fma(V6_CA, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
            n1 = fma(V6_CA, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
          } else {
            n0 = fma(V6_CW, kk[0], acc[0]);
            n2 = fma(V6_CW, kk[P - 2], acc[P - 2]);
            n1 = fma(V6_CW, kk[P - 1], acc[P - 1]);
            Qp += Qinc;                                   // stage 0 of the NEXT step
          }
          e_first[buf * T + t] = n0;
          *reinterpret_cast<double2*>(Qp + ((st + 1) & 3) * STG) = make_double2(n2, n1);
        }
#pragma unroll
        for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (st == 0) acc[i] = fma(V6_CW, kk[i], xd[i]);
          else if (st < 3) acc[i] = fma(V6_CW, kk[i], acc[i]);
          if (st < 3) s[i] = fma(V6_CA, kk[i], st == 2 ? xd[i] : xh[i]);
          else { s[i] = fma(V6_CW, kk[i], acc[i]); xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
        }
#undef V6_CA
#undef V6_CW
      }
      Lp += Linc;
    }
    // ---- registers -> staging -> coalesced 16-byte stores of the still-valid points [p, q)
#pragma unroll
    for (int i = 0; i < P; i += 2) *reinterpret_cast<double2*>(st_out + 18 * t + i) = make_double2(s[i], s[i + 1]);
    v6_group_sync<TG>(grp);
    {
      double* dst = a.out + (long long)(cur.x >> 1) * a.out_ld + cur.y + 2 * t;
      const double* srcs = st_out + 2 * t + 2 * (t >> 3);
      const int e_lo = cur.z - cur.y - 2 * t, e_hi = cur.w - cur.y - 2 * t;   // pair 2 T i is stored iff e_lo <= 2 T i < e_hi
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (2 * T * i >= e_lo && 2 * T * i < e_hi)
          *reinterpret_cast<double2*>(dst + 2 * T * i) = *reinterpret_cast<const double2*>(srcs + (2 * T + T / 4) * i);
      }
      if (emit) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");     // the columns (and, long since, the next tile)
        double* yb_row = a.emit.yb + (long long)(cur.x >> 1) * a.emit.ldy;
        for (int r = r_lo + t; r < r_hi; r += T) {
          const int c = s_cols[r - r_lo];
          const int e = c - cur.y;
          if (c >= cur.z && c < cur.w) yb_row[r] = st_out[e + 2 * (e >> 4)];
        }
      }
    }
    if (ti_next < 0) break;
    cur = nxt; ti = ti_next; te = te_next; r_lo = nr_lo; r_hi = nr_hi;
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool v6_shape(int variant, int* tg, int* g) {
  if (variant == 0) { *tg = 128; *g = 3; return true; }
  if (variant == 1) { *tg = 64; *g = 6; return true; }
  return false;
}

bool l96_v6_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj, int variant) {
  int tg, g;
  if (!v6_shape(variant, &tg, &g)) return false;
  if (traj || n_steps < 1 || n_steps > V6_MAX_STEPS || 12 * n_steps + 16 > v6_tile(tg)) return false;
  if ((n & 1) || (in_ld & 1) || (out_ld & 1) || (in_base & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (n < 4 || n + 12ll * n_steps >= (1ll << 30)) return false;
  return true;
}

// The flattened (member, position) space is cut into `ctas` equal contiguous ranges (even boundaries); inside a
// range, member by member, into chains: a fresh tile (first useful point `pos`, window from pos - 8 S, useful
// [pos, g0 + A)) followed by continuing tiles (window from g0, useful [g0, g0 + A)), A = W - 4 S rounded down to
// whole threads.  Chain lengths are guided: ceil(tiles left in the range / (2 G)), at least one.
bool l96_v6_plan(long long n, int k, int n_steps, int ctas, int variant, L96V6Plan* out) {
  int tg, G;
  if (!v6_shape(variant, &tg, &G)) return false;
  const int W = v6_tile(tg);
  if (n_steps < 1 || n_steps > V6_MAX_STEPS || 12 * n_steps + 16 > W || ctas < 1) return false;
  const int S = n_steps;
  const int A = ((W - 4 * S) / V6_P) * V6_P;
  const long long total = n * (long long)k;
  long long per = (ceil_div_ll(total, ctas) + 1) & ~1ll;                // points per CTA, even
  if (per < 2) per = 2;
  const int grid = (int)ceil_div_ll(total, per);
  std::vector<std::vector<int>> tiles((size_t)grid), chains((size_t)grid);
  size_t max_tiles = 1, max_chains = 1;
  for (int b = 0; b < grid; ++b) {
    const long long lo = (long long)b * per, hi = lo + per < total ? lo + per : total;
    std::vector<int>& tl = tiles[(size_t)b];
    std::vector<int>& ch = chains[(size_t)b];
    long long left_pts = hi - lo;
    for (long long f = lo; f < hi;) {                                    // one member segment at a time
      const int mem = (int)(f / n);
      long long pos = f - (long long)mem * n;
      const long long seg_end = ((long long)(mem + 1) * n < hi ? n : hi - (long long)mem * n);
      while (pos < seg_end) {
        const long long left_tiles = ceil_div_ll(left_pts, A);
        long long len = ceil_div_ll(left_tiles, 2 * G);
        if (len < 1) len = 1;
        ch.push_back((int)(tl.size() / 4));
        bool fresh = true;
        for (long long i = 0; i < len && pos < seg_end; ++i) {
          const long long g0 = fresh ? pos - 8 * S : pos;
          const long long q = g0 + A < seg_end ? g0 + A : seg_end;
          tl.push_back(2 * mem + (fresh ? 0 : 1)); tl.push_back((int)g0); tl.push_back((int)pos); tl.push_back((int)q);
          left_pts -= q - pos;
          pos = g0 + A;
          fresh = false;
        }
      }
      f = (long long)mem * n + seg_end;
    }
    ch.push_back((int)(tl.size() / 4));
    if (tl.size() / 4 > max_tiles) max_tiles = tl.size() / 4;
    if (ch.size() > max_chains) max_chains = ch.size();
  }
  out->grid = grid;
  out->variant = variant;
  out->tile_stride = (int)max_tiles;
  out->chain_stride = (int)max_chains + 1;
  out->hist_t = A / V6_P - 1;
  out->tiles.assign((size_t)grid * max_tiles * 4, 0);
  out->chains.assign((size_t)grid * (max_chains + 1), 0);
  for (int b = 0; b < grid; ++b) {
    std::copy(tiles[(size_t)b].begin(), tiles[(size_t)b].end(), out->tiles.begin() + (size_t)b * max_tiles * 4);
    int* c = out->chains.data() + (size_t)b * (max_chains + 1);
    c[0] = (int)chains[(size_t)b].size() - 1;
    std::copy(chains[(size_t)b].begin(), chains[(size_t)b].end(), c + 1);
  }
  return true;
}

template <int TG, int G>
static cudaError_t v6_launch_t(const L96V6Args& a, int grid, cudaStream_t st) {
  const size_t smem_max = (size_t)G * ((size_t)v6_group_doubles(TG) * sizeof(double) + (size_t)(v6_tile(TG) + 128) * sizeof(int));
  const size_t smem = a.emit.yb != nullptr ? smem_max : (size_t)G * (size_t)v6_group_doubles(TG) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(l96_v6_kernel<TG, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_chain(true, l96_v6_kernel<TG, G>, dim3((unsigned)grid), dim3(G * TG), smem, st, a);
}

cudaError_t l96_v6_launch(const double* in, double* out, const int* tiles_dev, const int* chains_dev, int grid,
                          int tile_stride, int chain_stride, int hist_t, int variant, long long n, int n_steps,
                          double dt, double F, long long in_ld, long long out_ld, bool periodic, long long in_base,
                          long long in_lo, long long in_hi, cudaStream_t st, const ObsEmit* emit) {
  L96V6Args a{};
  if (emit != nullptr) a.emit = *emit;
  a.in = in; a.out = out;
  a.in_ld = in_ld; a.out_ld = out_ld; a.in_base = in_base;
  a.in_lo = (int)in_lo; a.in_hi = (int)in_hi;
  a.n = (int)n; a.periodic = periodic ? 1 : 0;
  a.tiles = reinterpret_cast<const int4*>(tiles_dev); a.chains = chains_dev;
  a.tile_stride = tile_stride; a.chain_stride = chain_stride;
  a.n_steps = n_steps; a.hist_t = hist_t;
  a.dt = dt; a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  if (variant == 1) return v6_launch_t<64, 6>(a, grid, st);
  return v6_launch_t<128, 3>(a, grid, st);
}

}  // namespace dab
