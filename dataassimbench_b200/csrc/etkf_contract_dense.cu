// K3 on a COMPACT observation-space ensemble: C = Y'^T Rinv Y', g = Y'^T Rinv (y - ybar) from Yb[k][n_y] that the
// forecast kernels wrote while they stored their tiles (stationary networks: the observed columns of a tile are a
// contiguous run of location-ordered rows), instead of gathering the observed columns from the state.
//
// Replaces the same reference lines as etkf_contract.cu (dabench/dacycler/_etkf.py:82-92, 131, 137-146, 154 with the
// default H and R of dabench/dacycler/_dacycler.py:58-72).
//
// Why: at 30 % observation density the gather of contract_kernel moves 34 MB of 32-byte sectors for 10 MB of payload
// and every 32-row block of a CTA pays a dependent round trip (row table -> state) that its 64 DMMAs cannot hide
// (15.7 us at C3 for 4.4 us of tensor work).  With Yb compact, a CTA's rows are k contiguous runs: it fetches ALL of
// them with one wave of coalesced cp.async copies (one L2 round trip, every sector fully used), centres them in
// shared memory and contracts with DMMA.8x8x4 -- only the upper triangle of the symmetric C (warp w computes the
// tiles (w, w + o mod K8), o = 0 .. K8/2; the mirror image is written with the tile), g as one more DMMA per k-step
// with the innovation as a one-column B operand.  The partial results have the layout stats_reduce_kernel expects.
#include "common.cuh"
#include "kernels.h"

namespace dab {

namespace {

struct DenseArgs {
  const double* Yb;              // [k][ldy], row r of this analysis at column r
  long long ldy;
  int k;
  const double* obs_val;         // [n_rows] observed values of this analysis, in row order
  const double* obs_sw;          // nullable: 1 / sigma per row (diagonal R with unequal entries)
  long long n_rows;
  int rc;                        // row capacity of a CTA (multiple of 32)
  double* partial;               // [grid][KP * KP + KP]
};

template <int K8>
__global__ void __launch_bounds__(32 * K8, 1)
contract_dense_kernel(const DenseArgs a) {
  constexpr int KP = 8 * K8, NT = 32 * K8, O = K8 / 2 + 1;
  extern __shared__ __align__(16) double sm[];
  const int RC = a.rc, LD = RC + 4;               // (4 j + r) mod 16 distinct: conflict-free fragments
  double* Ys = sm;                                // [KP][LD] Yb^T rows of this CTA, then centred
  double* dv = Ys + (size_t)KP * LD;              // [RC] observed value, then innovation y - ybar
  double* yb = dv + RC;                           // [RC] row mean
  double* sw = yb + RC;                           // [RC] 1 / sigma (1 when R = sd^2 I: applied in the reduction)
  double* psum = sw + RC;                         // [K8][RC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fr = lane & 3;
  const int k = a.k;

  // contiguous, balanced row range of this CTA
  const long long per = a.n_rows / gridDim.x, extra = a.n_rows % gridDim.x;
  const long long r_begin = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const int nr = (int)(per + (blockIdx.x < extra ? 1 : 0));
  for (int r = tid; r < RC; r += NT) {
    dv[r] = r < nr ? a.obs_val[r_begin + r] : 0.0;
    sw[r] = (r < nr && a.obs_sw != nullptr) ? a.obs_sw[r_begin + r] : 1.0;
  }
  // everything above read the observation tables only; Yb is the previous kernel's output
  pdl_wait();
  pdl_launch_dependents();

  // ---- all rows of the CTA at once: warp w takes members w, w + K8, ..., lanes over rows (coalesced 8-byte copies)
  for (int r = lane; r < RC; r += 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = warp + i * K8;
      double* dst = Ys + (size_t)j * LD + r;
      if (r < nr && j < k) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                     "l"(a.Yb + (long long)j * a.ldy + r_begin + r) : "memory");
      } else {
        *dst = 0.0;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- row means over the members (per-warp partial sums, then across the warps), innovation
  double v[8];
  for (int r0 = 0; r0 < RC; r0 += 32) {          // RC / 32 passes; the values stay in registers for the centring
    const int r = r0 + lane;
    double ps = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = Ys[(size_t)(warp + i * K8) * LD + r]; ps += v[i]; }
    psum[warp * RC + r] = ps;
  }
  __syncthreads();
  {
    const double inv_k = 1.0 / (double)k;
    for (int r = lane; r < RC; r += 32) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < K8; ++w) tot += psum[w * RC + r];
      const double m = tot * inv_k;
      const bool live = r < nr;
      const double s = sw[r];
      if (live) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = warp + i * K8;
          if (j < k) { double* p = Ys + (size_t)j * LD + r; *p = (*p - m) * s; }
        }
      }
      if (warp == 0) dv[r] = live ? (dv[r] - m) * s : 0.0;
    }
  }
  __syncthreads();

  // ---- C tiles (warp, warp + o) and g strip `warp` over all rows, 4 rows per DMMA
  double c0[O], c1[O], g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int o = 0; o < O; ++o) { c0[o] = 0.0; c1[o] = 0.0; }
  const bool last_tile = warp < K8 / 2;           // the tile pair at distance K8/2 is computed by the lower warp only
  {
    const int steps = (nr + 3) >> 2;
    const double* arow = Ys + (size_t)(8 * warp + fm) * LD + fr;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
      const double av = arow[4 * s];
      dmma884(c0[0], c1[0], av, av);
#pragma unroll
      for (int o = 1; o < O; ++o) {
        if (o == O - 1 && !last_tile) break;
        const int t = (warp + o) & (K8 - 1);
        const double bv = Ys[(size_t)(8 * t + fm) * LD + 4 * s + fr];
        dmma884(c0[o], c1[o], av, bv);
      }
      const double dd = fm == 0 ? dv[4 * s + fr] : 0.0;
      dmma884(g0, g1, av, dd);
    }
  }
  // partial layout: [KP][KP] row-major then [KP]; every tile with its mirror image
  double* out = a.partial + (long long)blockIdx.x * (KP * KP + KP);
#pragma unroll
  for (int o = 0; o < O; ++o) {
    if (o == O - 1 && !last_tile) break;
    const int t = (warp + o) & (K8 - 1);
    const int row = 8 * warp + fm, col = 8 * t + 2 * fr;
    *reinterpret_cast<double2*>(out + row * KP + col) = make_double2(c0[o], c1[o]);
    if (o != 0) { out[col * KP + row] = c0[o]; out[(col + 1) * KP + row] = c1[o]; }
  }
  if (fr == 0) out[KP * KP + 8 * warp + fm] = g0;    // column 0 of the 8 x 8 result
}

size_t dense_smem_bytes(int k8, int rc) {
  return ((size_t)8 * k8 * (rc + 4) + 3 * (size_t)rc + (size_t)k8 * rc) * sizeof(double);
}

}  // namespace

// grid and row capacity, or false when a CTA's rows do not fit in shared memory (the gather kernel then runs)
bool contract_dense_plan(long long n_rows, int k, int sms, int* grid_out, int* rc_out) {
  if (n_rows < 1) return false;
  const int k8 = contract_k8(k);
  const int grid = contract_grid(n_rows, sms);
  const long long per = (n_rows + grid - 1) / grid;
  const int rc = (int)((per + 31) / 32) * 32;
  if (dense_smem_bytes(k8, rc) > 200 * 1024) return false;
  *grid_out = grid; *rc_out = rc;
  return true;
}

cudaError_t contract_dense_launch(const double* Yb, long long ldy, int k, const double* obs_val, const double* obs_sw,
                                  long long n_rows, int grid, int rc, double* partial, bool chain, cudaStream_t st) {
  DenseArgs a{};
  a.Yb = Yb; a.ldy = ldy; a.k = k; a.obs_val = obs_val; a.obs_sw = obs_sw; a.n_rows = n_rows; a.rc = rc; a.partial = partial;
  const int k8 = contract_k8(k);
  const size_t smem = dense_smem_bytes(k8, rc);
  cudaError_t e = cudaSuccess;
#define DAB_CD_CASE(K)                                                                                           \
  {                                                                                                              \
    static size_t set_for = 0;                                                                                   \
    if (smem > set_for) {                                                                                        \
      e = cudaFuncSetAttribute(contract_dense_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                            \
      set_for = smem;                                                                                            \
    }                                                                                                            \
    return launch_chain(chain, contract_dense_kernel<K>, dim3(grid), dim3(32 * K), smem, st, a);                 \
  }
  switch (k8) {
    case 2: DAB_CD_CASE(2)
    case 4: DAB_CD_CASE(4)
    case 8: DAB_CD_CASE(8)
    default: DAB_CD_CASE(16)
  }
#undef DAB_CD_CASE
}

}  // namespace dab
