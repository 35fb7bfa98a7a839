// K1, third design -- the CTA-tile forecast kernel of l96_forecast.cu made persistent.
//
// Replaces the same reference functions (ETKF._step_forecast dabench/dacycler/_etkf.py:64-80 ->
// Model.forecast -> Data.generate dabench/data/_data.py:86-211 -> integrate dabench/data/_utils.py:16-57
// with Lorenz96.rhs dabench/data/_lorenz96.py:119-151), classical RK4 for all members at once.
//
// What the measurements of round 2 say about the first design (profiles/forecast_probes_r02.md):
//   * its inner loop (tile of 128 x 16 points in registers, edge values through shared memory, one
//     barrier per RK stage, three independent CTAs per SM) keeps the FP64 pipe ~85 % busy;
//   * but a CTA spends 3-5 us loading its tile before it computes for ~8 us, the launch ends in a partial
//     wave (5.33 waves at C3), and the 240-point halo of a 20-step window is 11.7 % of a 2048-point tile.
//     Splitting the window into two LAUNCHES halves the halo and loses more than that to a second
//     load/store phase and a second tail (measured: 123 -> 129 us at C3).
// So: one persistent launch, 3 CTAs per SM; the window is cut into chunks (e.g. 2 x 10 steps: halo 5.9 %);
// a job = one tile of one member for one chunk; jobs are numbered chunk-major and handed out in that order
// by an atomic dispenser (a CTA that runs ahead simply takes more of them).  The next job's tile is prefetched with cp.async (L2 path) into a second staging buffer
// while the current one computes, so the load phase disappears; results leave through the staging
// buffer as coalesced 16-byte stores.  A job of chunk c > 0 needs the <= 3 jobs of chunk c-1 around it:
// one flag per job in global memory (thread 0: fence + store after the CTA's stores; thread 0 of the
// consumer polls before the prefetch).  Producers have lower job numbers, so they were claimed earlier, by
// CTAs that are running: the waits cannot deadlock.  The claimed job number reaches the threads through
// shared memory (a load at a uniform address): the loop stays uniform to the compiler, which keeps the RK4
// coefficients in uniform registers (see l96_forecast.cu).
//
// Requirements (l96_v3_eligible): even n and row strides, 16-byte aligned rows, no trajectory output.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "l96_dataflow.cuh"

namespace dab {

constexpr int V3_P = 16;                          // points per thread
constexpr int V3_T = 128;                         // threads per CTA
constexpr int V3_W = V3_P * V3_T;                 // points per tile
constexpr int V3_STAGE = V3_W + 2 * (V3_W / 16);  // padded staging doubles (16-byte pairs stay aligned, LDS.128 conflict-free)

// cp.async (L2 path) of a job's tile into a staging buffer: 1024 16-byte pairs, thread t takes pairs t, t + 128, ...
__device__ __forceinline__ void v3_prefetch(const L96V2Args& a, const V2Job& jb, double* stage, int t) {
  const double* row;
  int lo, hi;
  if (jb.c == 0) { row = a.in + (long long)jb.m * a.in_ld + a.in_base; lo = a.in_lo; hi = a.in_hi; }
  else {
    row = a.scratch[(jb.c - 1) & 1] + (long long)jb.m * a.scr_ld + a.scr_base;
    lo = a.ch[jb.c - 1].org; hi = a.ch[jb.c - 1].end;
  }
  if (a.periodic) { lo = 0; hi = a.n; }
  // pair index q = t + 128 i -> element e = 2 q = 2 t + 256 i; padded position e + 2 (e >> 4) = 2 t + 2 (t >> 3) + 288 i
  const unsigned dst0 = v2_smem(stage + 2 * t + 2 * (t >> 3));
  if (jb.g0 >= lo && jb.g0 + V3_W <= hi) {
    const double* src = row + jb.g0 + 2 * t;
#pragma unroll
    for (int i = 0; i < V3_W / (2 * V3_T); ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(src + 256 * i) : "memory");
  } else {
#pragma unroll 1
    for (int i = 0; i < V3_W / (2 * V3_T); ++i) {
      int g = jb.g0 + 2 * t + 256 * i;
      if (a.periodic) g = v2_wrap(g, a.n);
      else g = g < lo ? lo : (g > hi - 2 ? hi - 2 : g);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 288 * 8 * i), "l"(row + g) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// thread 0: are the producers of `jb` done?  (volatile loads served by L2)
__device__ __forceinline__ bool v3_deps_ready(const L96V2Args& a, const V2Job& jb) {
  if (jb.c == 0) return true;
  int j0, cnt;
  v2_deps_range(a, jb, j0, cnt);
  const int Jp = a.ch[jb.c - 1].J;
  const unsigned* f0 = a.flags + a.ch[jb.c - 1].job_base + (long long)jb.m * Jp;
  bool ok = true;
  for (int i = 0; i < cnt; ++i) ok = ok && (v2_flag(f0, j0, i, cnt, Jp) == a.epoch);
  return ok;
}
// the same without waiting for the answer: the (up to three) flag loads are issued here and looked at a
// stage of arithmetic later, so that thread 0 does not hold its warp -- and with it the CTA's next barrier --
// for an L2 round trip
// release of a job's flag by thread 0 (after a CTA barrier that followed the job's stores)
__device__ __forceinline__ void v3_release(const L96V2Args& a, int n) {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.flags + n), "r"(a.epoch) : "memory");
}
struct V3Deps { unsigned v0, v1, v2; bool many; };
__device__ __forceinline__ V3Deps v3_deps_load(const L96V2Args& a, const V2Job& jb) {
  V3Deps d; d.v0 = d.v1 = d.v2 = a.epoch; d.many = false;
  if (jb.c == 0) return d;
  int j0, cnt;
  v2_deps_range(a, jb, j0, cnt);
  const int Jp = a.ch[jb.c - 1].J;
  const unsigned* f0 = a.flags + a.ch[jb.c - 1].job_base + (long long)jb.m * Jp;
  d.many = cnt > 3;
  d.v0 = v2_flag(f0, j0, 0, cnt, Jp); d.v1 = v2_flag(f0, j0, 1, cnt, Jp); d.v2 = v2_flag(f0, j0, 2, cnt, Jp);
  return d;
}

__global__ void __launch_bounds__(V3_T, 3)
l96_v3_kernel(const L96V2Args a) {
  constexpr int P = V3_P, T = V3_T;
  extern __shared__ __align__(16) double smem[];
  double2* e_last2 = reinterpret_cast<double2*>(smem);  // [2][T] (s_{P-2}, s_{P-1}) of every thread
  double* e_first = smem + 4 * T;                       // [2][T] s_0 of every thread
  double* st_in = e_first + 2 * T;                      // prefetched tile of the next job
  double* st_out = st_in + V3_STAGE;                    // results on their way out
  __shared__ int s_ready, s_claim[2];
  const int t = threadIdx.x;
  const int total = a.total;
  const int tl = t > 0 ? t - 1 : 0;
  const int tr = t < T - 1 ? t + 1 : T - 1;
  pdl_wait();                 // the input is the previous kernel's output
  pdl_launch_dependents();
  // jobs are handed out in order by an atomic dispenser (thread 0 claims, the number travels through shared
  // memory: a shared load at a uniform address stays uniform for the compiler); a CTA holds two claims, the
  // job it computes and the one it prefetches
  // (thread 0 never waits for an L2 round trip at a point where the CTA's next barrier would wait with it:
  // the atomic is issued at the start of a job and its result is stored to shared memory at the end of it)
  auto claim_issue = [&]() -> unsigned long long { return t == 0 ? atomicAdd(a.counter, 1ull) : 0ull; };
  auto claim_store = [&](int slot, unsigned long long v) {
    if (t == 0) s_claim[slot] = v < (unsigned long long)total ? (int)v : total;
  };
  {
    const unsigned long long v0 = claim_issue(), v1 = claim_issue();
    claim_store(0, v0);
    claim_store(1, v1);
  }
  __syncthreads();
  // the last CTA to leave zeroes the dispenser for the next launch (counter[1] counts the leavers)
  auto leave = [&]() {
    if (t == 0 && atomicAdd(a.counter + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
      a.counter[0] = 0; a.counter[1] = 0;
      __threadfence();
    }
  };
  int n = s_claim[0];
  if (n >= total) { leave(); return; }
  V2Job cur = v2_decode(a, n, 0);
  if (t == 0) while (!v3_deps_ready(a, cur)) { }
  __syncthreads();
  v3_prefetch(a, cur, st_in, t);
  int slot = 1;                       // s_claim[slot] is the next job

  while (true) {
    // ---- this job's tile: staging -> registers (blocked: thread owns 16 consecutive points)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double s[P], xh[P], xd[P], acc[P];
#pragma unroll
    for (int i = 0; i < P; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(st_in + 18 * t + i);
      s[i] = v.x; s[i + 1] = v.y;
    }
    // ---- the next job of this CTA: thread 0 looks at its producers now, the prefetch follows a stage later
    const int nn = s_claim[slot];
    const bool have_next = nn < total;
    V2Job nxt = cur;
    V3Deps nd; nd.v0 = nd.v1 = nd.v2 = a.epoch; nd.many = false;
    if (have_next) {
      nxt = v2_decode(a, nn, cur.c);
      if (t == 0) nd = v3_deps_load(a, nxt);
    }
    const unsigned long long next_claim = have_next ? claim_issue() : 0ull;   // the job after next
    bool fetched = false;
#pragma unroll
    for (int i = 0; i < P; ++i) { xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
    int buf = 0;
    e_first[t] = s[0];
    e_last2[t] = make_double2(s[P - 2], s[P - 1]);

    const int ns = a.ch[cur.c].steps;
    for (int step = 0; step < ns; ++step) {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        __syncthreads();
        if (step == 0 && st == 1 && t == 0)                   // the flag loads had a stage to come back
          s_ready = (nd.v0 == a.epoch && nd.v1 == a.epoch && nd.v2 == a.epoch && !nd.many) ? 1 : 0;
        if (step == 0 && st == 2 && have_next && s_ready) {   // every thread left st_in long ago: refill it
          v3_prefetch(a, nxt, st_in, t);
          fetched = true;
        }
        const double2 L = e_last2[buf * T + tl];
        const double R = e_first[buf * T + tr];
        buf ^= 1;
#define V3_CA (st == 2 ? a.dt : a.h2)
#define V3_CW ((st == 0 || st == 3) ? a.h6 : a.h3)
        double kk[P];
        // edge points first so the next stage's edges can be published early
        kk[0] = fma(s[1] - L.x, L.y, -s[0]);
        kk[1] = fma(s[2] - L.y, s[0], -s[1]);
        kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
        kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
        if (st < 3) {
          const double n0 = fma(V3_CA, kk[0], st == 2 ? xd[0] : xh[0]);
          const double n2 = fma(V3_CA, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]);
          const double n1 = fma(V3_CA, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]);
          e_first[buf * T + t] = n0;
          e_last2[buf * T + t] = make_double2(n2, n1);
        } else {
          const double n0 = fma(V3_CW, kk[0], acc[0]);
          const double n2 = fma(V3_CW, kk[P - 2], acc[P - 2]);
          const double n1 = fma(V3_CW, kk[P - 1], acc[P - 1]);
          e_first[buf * T + t] = n0;
          e_last2[buf * T + t] = make_double2(n2, n1);
        }
#pragma unroll
        for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (st == 0) acc[i] = fma(V3_CW, kk[i], xd[i]);
          else if (st < 3) acc[i] = fma(V3_CW, kk[i], acc[i]);
          if (st < 3) s[i] = fma(V3_CA, kk[i], st == 2 ? xd[i] : xh[i]);
          else { s[i] = fma(V3_CW, kk[i], acc[i]); xh[i] = s[i] + a.h2F; xd[i] = s[i] + a.dtF; }
        }
#undef V3_CA
#undef V3_CW
      }
    }
    // ---- registers -> staging -> coalesced 16-byte stores of the still-valid points [p, q)
#pragma unroll
    for (int i = 0; i < P; i += 2) *reinterpret_cast<double2*>(st_out + 18 * t + i) = make_double2(s[i], s[i + 1]);
    __syncthreads();
    {
      const bool last = cur.c == a.n_chunks - 1;
      double* row = last ? a.out + (long long)cur.m * a.out_ld
                         : a.scratch[cur.c & 1] + (long long)cur.m * a.scr_ld + a.scr_base;
      double* dst = row + cur.g0 + 2 * t;
      const double* srcs = st_out + 2 * t + 2 * (t >> 3);
      const int e_lo = cur.p - cur.g0 - 2 * t, e_hi = cur.q - cur.g0 - 2 * t;   // pair 256 i is stored iff e_lo <= 256 i < e_hi
#pragma unroll
      for (int i = 0; i < V3_W / (2 * T); ++i) {
        if (256 * i >= e_lo && 256 * i < e_hi)
          *reinterpret_cast<double2*>(dst + 256 * i) = *reinterpret_cast<const double2*>(srcs + 288 * i);
      }
      if (last && a.emit.yb != nullptr)
        emit_observed(a.emit.loc, a.emit.row_start, a.emit.n_rows, a.emit.yb + (long long)cur.m * a.emit.ldy, st_out,
                      [](int e) { return e + 2 * (e >> 4); }, cur.g0, cur.p, cur.q, t, T);
    }
    __syncthreads();            // every thread's stores are issued (and st_out, the edge arrays are free again)
    if (t == 0) v3_release(a, n);   // the barrier above ordered the CTA's stores before thread 0's fence
    claim_store(slot ^ 1, next_claim);
    if (have_next && !fetched) {
      // the producers of the next job were not done when this one started (one of them may be this very
      // job, published just now): wait for them, fetch
      if (t == 0) while (!v3_deps_ready(a, nxt)) { }
      __syncthreads();
      v3_prefetch(a, nxt, st_in, t);
    }
    if (!have_next) break;
    cur = nxt;
    n = nn;
    slot ^= 1;
  }
  leave();
}

// ------------------------------------------------------------------------------------------ host side
int l96_v3_tile() { return V3_W; }

bool l96_v3_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj) {
  if (traj || n_steps < 1 || 12 * n_steps + 16 > V3_W) return false;
  if ((n & 1) || (in_ld & 1) || (out_ld & 1) || (in_base & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  if (n < 4 || n + 12ll * n_steps >= (1ll << 30)) return false;
  return true;
}

// scratch row stride for windows of `n_steps` steps over slabs of n points (extended layout, even)
long long l96_v3_scratch_ld(long long n, int n_steps) { return (n + 12ll * n_steps + 2) & ~1ll; }

// upper bound of the number of jobs (= flags): <= 8 chunks, each over at most n + 12 n_steps points in jobs of
// at least half a tile of useful points (fewer only when the whole row is shorter)
size_t l96_v3_flag_count(long long n, int k, int n_steps) {
  const long long per = (n + 12ll * n_steps) / (V3_W / 2) + 2;
  return (size_t)V2_MAX_CHUNKS * (size_t)k * (size_t)per;
}

cudaError_t l96_v3_launch(const double* in, double* out, double* scratch0, double* scratch1, long long scr_ld,
                          unsigned* flags, unsigned epoch, unsigned long long* counter, unsigned long long* ctr_base,
                          long long n, int k, int n_steps, int chunk_steps,
                          double dt, double F, long long in_ld, long long out_ld, bool periodic,
                          long long in_base, long long in_lo, long long in_hi, int sms, cudaStream_t st,
                          const ObsEmit* emit) {
  L96V2Args a{};
  if (emit != nullptr) a.emit = *emit;
  a.in = in; a.out = out; a.scratch[0] = scratch0; a.scratch[1] = scratch1;
  a.in_ld = in_ld; a.out_ld = out_ld; a.scr_ld = scr_ld;
  a.in_base = in_base; a.in_lo = (int)in_lo; a.in_hi = (int)in_hi;
  if (!l96_dataflow_plan(a, n, k, n_steps, chunk_steps, V3_W, periodic)) return cudaErrorInvalidValue;
  a.flags = flags; a.epoch = epoch;
  a.counter = counter; a.ctr_base = 0;
  a.dt = dt; a.h2 = 0.5 * dt; a.h3 = dt / 3.0; a.h6 = dt / 6.0; a.h2F = a.h2 * F; a.dtF = dt * F;
  const size_t smem = (6 * (size_t)V3_T + 2 * (size_t)V3_STAGE) * sizeof(double);
  static int per_sm = 0;                          // the grid must be co-resident: the flag waits rely on it
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(l96_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, l96_v3_kernel, V3_T, smem);
    if (e != cudaSuccess) return e;
    per_sm = occ < 1 ? 1 : occ;
  }
  long long grid = (long long)sms * per_sm;
  if (grid > a.total) grid = a.total;
  if (grid < 1) grid = 1;
  (void)ctr_base;                                // the kernel leaves the dispenser at zero
  return launch_chain(true, l96_v3_kernel, dim3((unsigned)grid), dim3(V3_T), smem, st, a);
}

}  // namespace dab
