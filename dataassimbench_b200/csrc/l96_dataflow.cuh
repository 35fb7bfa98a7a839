// Shared pieces of the dataflow forecast kernels (tools/experiments/l96_forecast_v2.cu: warp-private windows;
// l96_forecast_v3.cu: CTA tiles): a window of S steps is cut into chunks of a few steps; a job advances one
// tile of one member by one chunk; jobs are numbered chunk-major; one flag per job in global memory says
// "its output is visible" (== epoch of this launch).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace dab {

constexpr int V2_MAX_CHUNKS = 8;

struct V2Chunk {
  int steps;        // RK4 steps of this chunk
  int org, end;     // output region [org, end) in slab coordinates
  int U, J;         // useful points per job, jobs per member
  int job_base;     // number of the chunk's first job
  int pad0, pad1;
  double inv_J, inv_U;
};

struct L96V2Args {
  const double* in;
  double* out;
  double* scratch[2];
  long long in_ld, out_ld, scr_ld;    // member strides (doubles)
  long long in_base, scr_base;        // column of slab coordinate 0 in an input / scratch row
  int in_lo, in_hi;                   // EXT: valid slab-coordinate range of an input row
  int n;                              // PERIODIC: ring size; EXT: slab length
  int periodic, k, n_chunks, total;
  V2Chunk ch[V2_MAX_CHUNKS];
  unsigned* flags;                    // one per job; == epoch when the job's output is visible
  unsigned epoch;
  unsigned long long* counter;        // job dispenser: monotonic over launches, this launch starts at ctr_base
  unsigned long long ctr_base;
  double h2, h3, h6, dt, h2F, dtF;    // RK4 coefficients as uniform operands (see l96_forecast.cu)
  ObsEmit emit;                       // optional compact copy of the observed columns of the LAST chunk (kernels.h)
};

__device__ __forceinline__ unsigned v2_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct V2Job { int n, c, m, p, q, g0; };   // number, chunk, member, useful range [p, q), window start

// floor(x / d) for 0 <= x < 2^31 with inv = 1 / d: one FP64 multiply instead of an integer division chain
__device__ __forceinline__ int v2_div(int x, double inv) { return (int)(((double)x + 0.5) * inv); }

__device__ __forceinline__ V2Job v2_decode(const L96V2Args& a, int n, int c_hint) {
  V2Job jb;
  int c = c_hint;                               // jobs arrive in increasing order: the chunk only moves forward
  while (c + 1 < a.n_chunks && n >= a.ch[c + 1].job_base) ++c;
  const V2Chunk& ch = a.ch[c];
  const int r = n - ch.job_base;
  jb.n = n; jb.c = c;
  jb.m = v2_div(r, ch.inv_J);
  const int j = r - jb.m * ch.J;
  jb.p = ch.org + j * ch.U;
  jb.q = jb.p + ch.U < ch.end ? jb.p + ch.U : ch.end;
  jb.g0 = jb.p - 8 * ch.steps;
  return jb;
}

__device__ __forceinline__ int v2_wrap(int g, int n) {      // ring coordinate of g (any distance from [0, n))
  if ((unsigned)g < (unsigned)n) return g;
  if (g < 0) { g += n; if (g >= 0) return g; }
  else { g -= n; if (g < n) return g; }
  g %= n;
  return g < 0 ? g + n : g;
}

// The jobs of chunk c-1 a job depends on.  RAW: their output intersects my input window; WAR: their
// input window intersects the range I overwrite (same buffer, two chunks apart).  Both are covered by
// the jobs whose output range meets [p - 8 s, q + 8 s), s the larger of the two chunks' steps: <= 3 jobs
// unless the slab is much shorter than a tile.
struct V2Deps { unsigned v0, v1, v2; bool many; };

__device__ __forceinline__ void v2_deps_range(const L96V2Args& a, const V2Job& jb, int& j0, int& cnt) {
  const V2Chunk& cp = a.ch[jb.c - 1];
  const int s0 = a.ch[jb.c].steps;
  const int sm = s0 > cp.steps ? s0 : cp.steps;
  const int lo = jb.p - 8 * sm - cp.org, hi = jb.q + 8 * sm - 1 - cp.org;
  if (a.periodic) {                      // ring: job of the wrapped coordinate, then upwards around the ring
    if (hi - lo + 1 >= a.n) { j0 = 0; cnt = cp.J; }
    else {
      j0 = v2_div(v2_wrap(lo, a.n), cp.inv_U);
      cnt = v2_div(v2_wrap(hi, a.n), cp.inv_U) - j0;
      if (cnt < 0) cnt += cp.J;
      cnt += 1;
    }
  } else {
    j0 = lo >= 0 ? v2_div(lo, cp.inv_U) : 0;
    int j1 = hi >= 0 ? v2_div(hi, cp.inv_U) : -1;
    if (j1 > cp.J - 1) j1 = cp.J - 1;
    cnt = j1 - j0 + 1;
  }
}

__device__ __forceinline__ unsigned v2_flag(const unsigned* f0, int j0, int i, int cnt, int Jp) {
  int jj = j0 + (i < cnt ? i : 0);
  if (jj >= Jp) jj -= Jp;
  return *reinterpret_cast<const volatile unsigned*>(f0 + jj);
}


// Host side: cut `n_steps` into chunks of <= chunk_steps and each chunk's output region into jobs of
// <= tile - 12 steps useful points (even, balanced).  Returns false if it does not fit.
inline bool l96_dataflow_plan(L96V2Args& a, long long n, int k, int n_steps, int chunk_steps, int tile, bool periodic) {
  if (chunk_steps < 1) chunk_steps = 1;
  const int nc = (n_steps + chunk_steps - 1) / chunk_steps;
  if (nc < 1 || nc > V2_MAX_CHUNKS) return false;
  a.n = (int)n; a.periodic = periodic ? 1 : 0; a.k = k; a.n_chunks = nc;
  a.scr_base = periodic ? 0 : 8ll * n_steps;
  int done = 0;
  long long base = 0;
  for (int c = 0; c < nc; ++c) {
    const int sc = (n_steps - done + (nc - c) - 1) / (nc - c);      // near-equal chunks
    done += sc;
    const int rem = n_steps - done;
    V2Chunk& ch = a.ch[c];
    ch.steps = sc;
    ch.org = periodic ? 0 : -8 * rem;
    const long long len = periodic ? n : n + 12ll * rem;
    ch.end = ch.org + (int)len;
    const int cap = tile - 12 * sc;
    if (cap < 16) return false;
    long long J = (len + cap - 1) / cap;
    long long U = (len + J - 1) / J;
    U = (U + 1) & ~1ll;                                             // even: 16-byte pairs never straddle jobs
    if (U > cap) U = cap & ~1;
    J = (len + U - 1) / U;
    ch.U = (int)U; ch.J = (int)J;
    ch.inv_U = 1.0 / (double)U; ch.inv_J = 1.0 / (double)J;
    ch.job_base = (int)base;
    base += J * k;
    if (base >= (1ll << 30)) return false;
  }
  a.total = (int)base;
  return true;
}

}  // namespace dab
