// EXPERIMENT (not part of the product build).  Correct (all GPU tests passed with it wired into dab_cycler_stats and
// dab_etkf_stats_f64), but not faster where it matters: at C3 its critical path is a chain of latency-bound
// steps -- row tables, gather (every load in flight at once), centring (two more barriers), 204 DMMAs per warp,
// partial store + fence, grid barrier, slice reduction -- that measured 20.9 us under ncu (tensor pipe 18.5 % of
// active cycles) against 15.7 + 5.3 us for contract_kernel + stats_reduce_kernel, and the whole cycle got SLOWER
// (182.3 vs 175.5 us, 5485 vs 5697 cycles/s) because the two short kernels overlap their neighbours under
// programmatic dependent launch while a kernel with a grid barrier cannot start its tail early.
// profiles/analysis_probes_r02.md has the numbers.
//
// K3, one-launch form for observation sets that fit in shared memory: gather + centring + contraction +
// reduction (+ the multi-rank push) in ONE kernel.
//
// Replaces the same reference lines as etkf_contract.cu (dabench/dacycler/_etkf.py:82-92, 131, 137-146, 154 with
// the default H and R of dabench/dacycler/_dacycler.py:58-72):  C = Y'^T Rinv Y',  g = Y'^T Rinv (y - ybar).
//
// Why a second form.  At C3 (19 661 rows, k = 64) a CTA of the pipelined kernel owns ~133 rows = 5 blocks of 32 and
// every block pays a gather round trip that its 64 DMMAs cannot hide (15.7 us for 4.4 us of tensor work), and the
// 148 partial results then need a second kernel (5-8 us, launch-bound).  Here
//   * a CTA gathers ALL its rows at once -- k x rows doubles straight into shared memory, every load of the CTA in
//     flight together, so the L2 latency is paid once;
//   * centres them in place (row mean over the members) and contracts with DMMA.8x8x4; C is symmetric, so warp w
//     computes only the 8 x 8 tiles (w, w + o mod K8), o = 0 .. K8/2 (36 of 64 at k = 64), and g is one more
//     DMMA per k-step with the innovation as a one-column B operand instead of a serial FMA chain;
//   * the CTAs meet at a grid barrier (one CTA per SM: all co-resident; a monotonic counter in global memory) and
//     each then sums its slice of the elements over all partials in a fixed order (bit-reproducible, no atomics on
//     data), applies 1/sigma^2 and writes both triangles of C -- or, on a grid-sharded run, stores the values into
//     slot `rank` of every rank's exchange buffer over NVLink and the last CTA raises the flags, exactly as
//     stats_reduce_kernel does.
#include "../../dataassimbench_b200/csrc/common.cuh"
#include "../../dataassimbench_b200/csrc/kernels.h"

namespace dab {

namespace {

struct FusedArgs {
  const double* X;
  long long ld;
  int k;
  const int* obs_loc;
  const double* obs_val;
  const long long* seg;
  int n_seg;
  long long n_rows;
  int rc;                        // row capacity of a CTA (multiple of 32)
  double* partial;               // [grid][P]
  double r_inv;
  double* stats;                 // world <= 1
  StatsPush push;
  unsigned long long* bar;       // grid barrier: monotonic arrival counter
  unsigned long long bar_target; // value it has when every CTA of this launch has arrived
};

__device__ __forceinline__ void fused_fetch_row(const FusedArgs& a, long long r, int& loc, double& val) {
  int s = 0;
  while (s + 1 < a.n_seg && a.seg[3 * (s + 1) + 2] <= r) ++s;
  const long long p = a.seg[3 * s] + (r - a.seg[3 * s + 2]);
  loc = a.obs_loc[p]; val = a.obs_val[p];
}

template <int K8>
__global__ void __launch_bounds__(32 * K8, 1)
contract_fused_kernel(const FusedArgs a) {
  constexpr int KP = 8 * K8, NT = 32 * K8, O = K8 / 2 + 1;
  constexpr int P = K8 * O * 64 + KP;             // doubles per partial: tiles (w, o) then g
  extern __shared__ __align__(16) double sm[];
  const int RC = a.rc, LD = RC + 4;               // (4 j + r) mod 16 distinct: conflict-free fragments
  double* Ys = sm;                                // [KP][LD] gathered, then centred, Yb^T
  double* dv = Ys + (size_t)KP * LD;              // [RC] observed value, then innovation y - ybar
  double* yb = dv + RC;                           // [RC] row mean
  int* s_loc = reinterpret_cast<int*>(yb + RC);   // [RC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fr = lane & 3;
  const int k = a.k;

  // contiguous, balanced row range of this CTA
  const long long per = a.n_rows / gridDim.x, extra = a.n_rows % gridDim.x;
  const long long r_begin = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const int nr = (int)(per + (blockIdx.x < extra ? 1 : 0));
  for (int r = tid; r < RC; r += NT) {
    int loc = -1; double val = 0.0;
    if (r < nr) fused_fetch_row(a, r_begin + r, loc, val);
    s_loc[r] = loc; dv[r] = val;
  }
  __syncthreads();
  // everything above read the observation tables only; the state is the previous kernel's output
  pdl_wait();
  pdl_launch_dependents();

  // ---- gather: warp w takes members w, w + K8, ..., lanes over rows; every element is one 8-byte cp.async, so all
  // loads of the CTA are in flight together and the L2 round trip is paid once
  for (int r = lane; r < RC; r += 32) {
    const int loc = s_loc[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      double* dst = Ys + (size_t)j * LD + r;
      if (loc >= 0 && j < k) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                     "l"(a.X + (long long)j * a.ld + loc) : "memory");
      } else {
        *dst = 0.0;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- row means over the members (per-warp partial sums, then across the warps), innovation
  double* psum = reinterpret_cast<double*>(s_loc + RC);   // [K8][RC] (RC is a multiple of 32: 8-byte aligned)
  for (int r = lane; r < RC; r += 32) {
    double ps = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) ps += Ys[(size_t)(warp + i * K8) * LD + r];
    psum[warp * RC + r] = ps;
  }
  __syncthreads();
  {
    const double inv_k = 1.0 / (double)k;
    for (int r = tid; r < RC; r += NT) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < K8; ++w) tot += psum[w * RC + r];
      const double m = tot * inv_k;
      const bool live = s_loc[r] >= 0;
      yb[r] = live ? m : 0.0;
      dv[r] = live ? dv[r] - m : 0.0;
    }
  }
  __syncthreads();
  // ---- centre in place (padding members j >= k and padding rows stay zero)
  for (int r = lane; r < RC; r += 32) {
    const double m = yb[r];
    if (s_loc[r] >= 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = warp + i * K8;
        if (j < k) Ys[(size_t)j * LD + r] -= m;
      }
    }
  }
  __syncthreads();

  // ---- C tiles (warp, warp + o) and g strip `warp` over all rows, 4 rows per DMMA
  double c0[O], c1[O], g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int o = 0; o < O; ++o) { c0[o] = 0.0; c1[o] = 0.0; }
  {
    const int steps = (nr + 3) >> 2;
    const double* arow = Ys + (size_t)(8 * warp + fm) * LD + fr;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
      const double av = arow[4 * s];
      dmma884(c0[0], c1[0], av, av);
#pragma unroll
      for (int o = 1; o < O; ++o) {
        const int t = (warp + o) & (K8 - 1);
        const double bv = Ys[(size_t)(8 * t + fm) * LD + 4 * s + fr];
        dmma884(c0[o], c1[o], av, bv);
      }
      const double dd = fm == 0 ? dv[4 * s + fr] : 0.0;
      dmma884(g0, g1, av, dd);
    }
  }
  {
    double* out = a.partial + (size_t)blockIdx.x * P;
#pragma unroll
    for (int o = 0; o < O; ++o)
      *reinterpret_cast<double2*>(out + (size_t)(warp * O + o) * 64 + 2 * lane) = make_double2(c0[o], c1[o]);
    if (fr == 0) out[K8 * O * 64 + 8 * warp + fm] = g0;      // column 0 of the 8 x 8 result
  }

  // ---- grid barrier: every partial is in global memory
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned long long t0, t1, seen = atomicAdd(a.bar, 1ull) + 1ull;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (seen < a.bar_target) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.bar) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) __trap();         // a CTA of this launch never ran: do not hang the GPU
    }
  }
  __syncthreads();

  // ---- this CTA's slice of the elements, summed over all partials in a fixed order
  const int n_part = gridDim.x;
  const int e_lo = (int)((long long)P * blockIdx.x / gridDim.x), e_hi = (int)((long long)P * (blockIdx.x + 1) / gridDim.x);
  const int world = a.push.world;
  for (int e = e_lo + warp; e < e_hi; e += K8) {
    int i, j, mirror = 0;
    if (e < K8 * O * 64) {
      const int ti = e >> 6, within = e & 63;
      const int w = ti / O, o = ti - w * O, t = (w + o) & (K8 - 1);
      if (o == K8 / 2 && w >= K8 / 2) continue;      // the same tile pair is taken from warp w - K8/2
      const int l = within >> 1;
      i = 8 * w + (l >> 2); j = 8 * t + 2 * (l & 3) + (within & 1);
      mirror = o != 0;
    } else {
      i = -1; j = e - K8 * O * 64;
    }
    double s = 0.0;
    for (int p = lane; p < n_part; p += 32) s += __ldcg(a.partial + (size_t)p * P + e);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    s *= a.r_inv;
    if (lane == 0 && j < k && i < k) {
      const int d0 = i < 0 ? k * k + j : i * k + j;
      const int d1 = (i >= 0 && mirror) ? j * k + i : -1;
      if (world <= 1) {
        a.stats[d0] = s;
        if (d1 >= 0) a.stats[d1] = s;
      } else {
        for (int g = 0; g < world; ++g) {
          a.push.slot[g][d0] = s;
          if (d1 >= 0) a.push.slot[g][d1] = s;
        }
      }
    }
  }
  if (world > 1) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int prev = atomicAdd(a.push.done_ctr, 1u);
      if (prev == gridDim.x - 1) {               // every CTA of this rank has pushed (and fenced)
        *a.push.done_ctr = 0;
        __threadfence_system();
        for (int g = 0; g < world; ++g) *reinterpret_cast<volatile unsigned long long*>(a.push.flag[g]) = a.push.seq;
      }
    }
  }
}

size_t fused_smem_bytes(int k8, int rc) {
  // Ys [8 k8][rc + 4], dv [rc], yb [rc], s_loc [rc] (int), psum [k8][rc]
  return ((size_t)8 * k8 * (rc + 4) + 2 * (size_t)rc + (size_t)k8 * rc) * sizeof(double) + (size_t)rc * sizeof(int);
}

}  // namespace

// grid and row capacity of the one-launch form, or false when a CTA's rows do not fit in shared memory
bool contract_fused_plan(long long n_rows, int k, int sms, int* grid_out, int* rc_out) {
  if (n_rows < 1) return false;
  const int k8 = contract_k8(k);
  long long g = (n_rows + 31) / 32;
  if (g > sms) g = sms;
  const long long per = (n_rows + g - 1) / g;
  const int rc = (int)((per + 31) / 32) * 32;
  if (fused_smem_bytes(k8, rc) > 200 * 1024) return false;
  *grid_out = (int)g; *rc_out = rc;
  return true;
}

size_t contract_fused_partial_doubles(int k, int grid) {
  const int k8 = contract_k8(k);
  return (size_t)grid * ((size_t)k8 * (k8 / 2 + 1) * 64 + 8 * k8);
}

cudaError_t contract_fused_launch(const double* X, long long ld, int k, const int* obs_loc, const double* obs_val,
                                  const long long* seg, int n_seg, long long n_rows, int grid, int rc,
                                  double* partial, double r_inv, double* stats, const StatsPush* push,
                                  unsigned long long* bar, unsigned long long bar_target, bool chain, cudaStream_t st) {
  FusedArgs a{};
  a.X = X; a.ld = ld; a.k = k; a.obs_loc = obs_loc; a.obs_val = obs_val; a.seg = seg; a.n_seg = n_seg;
  a.n_rows = n_rows; a.rc = rc; a.partial = partial; a.r_inv = r_inv; a.stats = stats;
  if (push) a.push = *push; else a.push.world = 1;
  a.bar = bar; a.bar_target = bar_target;
  const int k8 = contract_k8(k);
  const size_t smem = fused_smem_bytes(k8, rc);
  cudaError_t e = cudaSuccess;
#define DAB_CF_CASE(K)                                                                                           \
  {                                                                                                              \
    static size_t set_for = 0;                                                                                   \
    if (smem > set_for) {                                                                                        \
      e = cudaFuncSetAttribute(contract_fused_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                            \
      set_for = smem;                                                                                            \
    }                                                                                                            \
    return launch_chain(chain, contract_fused_kernel<K>, dim3(grid), dim3(32 * K), smem, st, a);                 \
  }
  switch (k8) {
    case 2: DAB_CF_CASE(2)
    case 4: DAB_CF_CASE(4)
    case 8: DAB_CF_CASE(8)
    default: DAB_CF_CASE(16)
  }
#undef DAB_CF_CASE
}

}  // namespace dab
