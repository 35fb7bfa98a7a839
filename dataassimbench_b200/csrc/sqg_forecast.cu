// Surface-QG turbulence ensemble forecast behind the forecast plugin point (second half).
//
// Replaces, for all members at once, SQGTurb.integrate -> lax.scan(SQGTurb._rk4) -> SQGTurb.rhs
// (dabench/data/_sqgturb.py:441-504, :345-364, :506-548) with its helpers _invert (:258-270), _specpad (:293-309),
// _xyderiv_dealias (:334-343), _spectrunc (:311-325) and the final `ifft2(values)` (:496).
//
// Design.  The state is tiny (2 x N x (N/2+1) complex, N = 96 by default) and every RK stage needs ten 2-D real FFTs
// on the 3N/2 dealiasing grid of which only N x (N/2+1) inputs are non-zero and only N x N/2 outputs are kept.  At
// that size a radix FFT is latency- and launch-bound, while the pruned transforms written as products with constant
// DFT matrices are FP64 GEMMs with all members batched into the long dimension -- work the FP64 pipe does at its
// peak rate.  Per RK stage:
//   prep      spectral state -> the 8 fields i k psi, i l psi, i k q, i l q (two levels each) on the non-zero block
//             of the padded spectrum, including the 2.25 factor and the conjugated Nyquist column       (elementwise)
//   gemm b    inverse transform along y:  [2M x 2N] constant  x  [2N x members 8 (N/2+1)]
//   gemm c    inverse c2r transform along x, per (member, field):  [M x (N/2+1)] x [(N/2+1) x M], real and imaginary
//             parts as two accumulated products
//   jacobian  psi_x q_y - psi_y q_x on the M x M grid                                                    (elementwise)
//   gemm d    forward r2c along x, truncated to N/2 columns:  [members 2 M x M] x [M x N]
//   gemm e    forward transform along y, truncated to N rows:  [2N x M] x [M x N/2], two accumulated products
//   update    tendency (thermal relaxation, Ekman terms, -J), RK4 bookkeeping in the reference's order, hyperdiffusion
//             factor after the fourth stage, and the next stage's prep in the same kernel                  (elementwise)
// All members of an ensemble share every launch.  The GEMM is one kernel on the FP64 tensor pipe (DMMA.8x8x4, see
// sqg_gemm_kernel): a first version with 4 x 4 outputs per thread on DFMA was shared-memory-bound at 20 % of peak
// (1 LDS.128 per 4 DFMAs).  M = 3N/2 and every table come from the caller (the Python mirror builds them the way the reference's
// __init__ does, including its float32 / complex64 casts), so the kernels hold no physics constants.
#include <cmath>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace dab {

namespace {

// ------------------------------------------------------------------------------------------ batched GEMM
// C[b] = A1[b] B1[b] (+ A2[b] B2[b]); every operand row-major with unit column stride, its own row and batch stride
struct GemmOp { const double* p; long long rs, bs; };
struct GemmArgs {
  GemmOp A[2], B[2];
  double* C; long long c_rs, c_bs, c_cs;      // C alone may have a column stride (interleaved complex output)
  int m, n, K, nseg;
};

// One CTA = 4 warps (2 x 2) on a (16 WT) x (16 WT) tile of C, k-tiles of 16; a warp owns WT x WT DMMA.8x8x4 tiles
// (WT A fragments + WT B fragments per k-step for WT^2 DMMAs).  A is staged as [row][k] with row stride 20, B as
// [k][column] with row stride 16 WT + 4: both = 4 mod 16, which spreads the 32 8-byte words of a fragment load over
// all banks (no conflicts), and the staging stores are contiguous per half-warp.  The next k-tile is fetched into
// registers while the current one is multiplied.  WT is picked per product so that the padded tile grid wastes the
// least work (N = 96: every dimension but the long batched one is a multiple of 48).
constexpr int GK = 16, LDA = GK + 4;

template <int WT>
__global__ void __launch_bounds__(128, WT >= 4 ? 3 : 4)
sqg_gemm_kernel(const GemmArgs g) {
  constexpr int BT = 16 * WT, LDB = BT + 4, NP = 2 * WT;     // NP = BT * GK / 128 elements per thread and operand
  __shared__ __align__(16) double As[BT * LDA];
  __shared__ __align__(16) double Bs[GK * LDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fr = lane & 3;
  const int m_base = (warp >> 1) * 8 * WT, n_base = (warp & 1) * 8 * WT;
  const int i0 = blockIdx.y * BT, j0 = blockIdx.x * BT;
  const long long b = blockIdx.z;
  double acc[WT][WT][2];
#pragma unroll
  for (int i = 0; i < WT; ++i)
#pragma unroll
    for (int j = 0; j < WT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const int kt = (g.K + GK - 1) / GK;
  const int total = kt * g.nseg;
  double ra[NP], rb[NP];
  // Loader geometry, fixed for the whole kernel: element p of this thread is A[row (tid >> 4) + 8 p][k (tid & 15)] and
  // B[k (tid + 128 p) / BT][column (tid + 128 p) % BT] of the tile.  Everything that does not change from k-tile to
  // k-tile -- base pointers of the segment, the B offsets, which rows / columns exist -- is computed once (the first
  // version recomputed it per element and spent 13 integer instructions per load: IMAD 2.5x the DMMA count).
  unsigned a_ok = 0, b_ok = 0;
  int a_off[NP], b_off[NP];                  // element offsets inside a k-tile (32-bit: the host checks 64 rs < 2^31)
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int idx = tid + 128 * p;
    if (i0 + (idx >> 4) < g.m) a_ok |= 1u << p;
    if (j0 + idx % BT < g.n) b_ok |= 1u << p;
  }
  const bool interior = i0 + BT <= g.m && j0 + BT <= g.n;
  const int a_kk = tid & 15;
  const double* Ab = nullptr;
  const double* Bb = nullptr;
  long long b_rs = 0;
  auto segment = [&](int seg) {
    const int ars = (int)g.A[seg].rs, brs = (int)g.B[seg].rs;
    Ab = g.A[seg].p + b * g.A[seg].bs + (long long)(i0 + (tid >> 4)) * g.A[seg].rs + a_kk;
    Bb = g.B[seg].p + b * g.B[seg].bs + j0;
    b_rs = g.B[seg].rs;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int idx = tid + 128 * p;
      a_off[p] = 8 * p * ars;
      b_off[p] = (idx / BT) * brs + idx % BT;
    }
  };
  auto fetch = [&](int it) {
    if (it == 0) segment(0);
    else if (it == kt) segment(1);
    const int k0 = (it >= kt ? it - kt : it) * GK;
    const double* A = Ab + k0;
    const double* B = Bb + (long long)k0 * b_rs;
    if (interior && k0 + GK <= g.K) {                      // CTA-uniform: no predicates, no branches around loads
#pragma unroll
      for (int p = 0; p < NP; ++p) { ra[p] = A[a_off[p]]; rb[p] = B[b_off[p]]; }
    } else if (k0 + GK <= g.K) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        ra[p] = (a_ok >> p & 1u) ? A[a_off[p]] : 0.0;
        rb[p] = (b_ok >> p & 1u) ? B[b_off[p]] : 0.0;
      }
    } else {                                               // ragged last k-tile
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        ra[p] = ((a_ok >> p & 1u) && k0 + a_kk < g.K) ? A[a_off[p]] : 0.0;
        rb[p] = ((b_ok >> p & 1u) && k0 + (tid + 128 * p) / BT < g.K) ? B[b_off[p]] : 0.0;
      }
    }
  };
  const int last_rem = g.K - (kt - 1) * GK;
  const int last_ksteps = last_rem >= GK ? GK / 4 : (last_rem + 3) >> 2;
  fetch(0);
  for (int it = 0; it < total; ++it) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int idx = tid + 128 * p;
      As[(idx >> 4) * LDA + (idx & 15)] = ra[p];
      Bs[(idx / BT) * LDB + idx % BT] = rb[p];
    }
    __syncthreads();
    // a ragged last k-tile (of either segment) only pays for its own k-steps
    const int ksteps = (it == kt - 1 || it == total - 1) ? last_ksteps : GK / 4;
    if (it + 1 < total) fetch(it + 1);
    for (int ks = 0; ks < ksteps; ++ks) {
      double av[WT], bv[WT];
#pragma unroll
      for (int i = 0; i < WT; ++i) av[i] = As[(m_base + 8 * i + fm) * LDA + 4 * ks + fr];
#pragma unroll
      for (int j = 0; j < WT; ++j) bv[j] = Bs[(4 * ks + fr) * LDB + n_base + 8 * j + fm];
#pragma unroll
      for (int i = 0; i < WT; ++i)
#pragma unroll
        for (int j = 0; j < WT; ++j) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
    }
  }
  double* C = g.C + b * g.c_bs;
#pragma unroll
  for (int i = 0; i < WT; ++i) {
    const int r = i0 + m_base + 8 * i + fm;
    if (r >= g.m) continue;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
      const int c = j0 + n_base + 8 * j + 2 * fr;
      if (c < g.n) C[(long long)r * g.c_rs + c * g.c_cs] = acc[i][j][0];
      if (c + 1 < g.n) C[(long long)r * g.c_rs + (c + 1) * g.c_cs] = acc[i][j][1];
    }
  }
}

cudaError_t gemm(cudaStream_t st, int batch, int m, int n, int K, GemmOp A1, GemmOp B1, double* C, long long c_rs,
                 long long c_bs, const GemmOp* A2 = nullptr, const GemmOp* B2 = nullptr, long long c_cs = 1) {
  GemmArgs g{};
  g.A[0] = A1; g.B[0] = B1; g.nseg = 1;
  if (A2 != nullptr) { g.A[1] = *A2; g.B[1] = *B2; g.nseg = 2; }
  g.C = C; g.c_rs = c_rs; g.c_bs = c_bs; g.c_cs = c_cs; g.m = m; g.n = n; g.K = K;
  for (int sgi = 0; sgi < g.nseg; ++sgi)
    if (g.A[sgi].rs >= (1ll << 25) || g.B[sgi].rs >= (1ll << 25)) return cudaErrorInvalidValue;   // 32-bit tile offsets
  // tile edge: the padded work shared by the SMs -- or, when there are fewer CTAs than that takes, one CTA's own
  // tile (small ensembles: a 48-edge grid of 64 CTAs leaves more than half of the GPU idle) -- weighted by the
  // fragment loads a DMMA costs at that warp tile (1, 2/3, 1/2 per DMMA for 32 / 48 / 64: measured on the y-inverse
  // product, 48 beats 32 by 12 % at equal padding)
  int best = 64;
  double best_cost = -1.0;
  for (int bt : {64, 48, 32}) {
    const double padded = (double)((m + bt - 1) / bt) * bt * (double)((n + bt - 1) / bt) * bt * (double)batch;
    const double per_sm = padded / 148.0, one_cta = (double)bt * bt;
    const double w = (per_sm > one_cta ? per_sm : one_cta) * (1.0 + 8.0 / bt);
    if (best_cost < 0.0 || w < best_cost) { best = bt; best_cost = w; }
  }
  dim3 grid((unsigned)((n + best - 1) / best), (unsigned)((m + best - 1) / best), (unsigned)batch);
  if (best == 64) sqg_gemm_kernel<4><<<grid, 128, 0, st>>>(g);
  else if (best == 48) sqg_gemm_kernel<3><<<grid, 128, 0, st>>>(g);
  else sqg_gemm_kernel<2><<<grid, 128, 0, st>>>(g);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ elementwise kernels
struct SqgDev {
  int N, NK, M, nm;                  // grid, N/2+1, 3N/2, members
  int ekman, symmetric;
  double dt, inv_tdiab, r;
  const double *kx, *ly;             // [NK], [N]: imaginary parts of ik_pad / il_pad on the non-zero block
  const double *hovermu, *tanhmu, *sinhmu, *ksqlsq, *hyperdiff;   // [N][NK]
  const double2* pveq;               // [2][N][NK]
  double2 *x, *xs, *acc;             // state, stage state, RK accumulator: [nm][2][N][NK]
  double* spec8;                     // [2 (re, im)][N][nm][8][NK]
  const double* jst;                 // [nm][2][2N][N/2]
  double* st2;                       // nullable: [2][N][nm][2][NK] new state for the trajectory transform
};

__device__ __forceinline__ double2 cmul_r(double a, double2 z) { return make_double2(a * z.x, a * z.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cdiv_r(double2 z, double a) { return make_double2(z.x / a, z.y / a); }

// _invert, _sqgturb.py:258-270, at one wavenumber
__device__ __forceinline__ void sqg_invert(const SqgDev& d, int rc, double2 q0, double2 q1, double2* p0, double2* p1) {
  const double hov = d.hovermu[rc], th = d.tanhmu[rc], sh = d.sinhmu[rc];
  *p0 = cmul_r(hov, csub(cdiv_r(q1, sh), cdiv_r(q0, th)));
  *p1 = cmul_r(hov, csub(cdiv_r(q1, th), cdiv_r(q0, sh)));
}

// the non-zero block of ik_pad * specpad(z), il_pad * specpad(z) at (r, c): 2.25 z (conjugated in the Nyquist
// column, _sqgturb.py:304-308) times i kx / i ly
__device__ __forceinline__ void sqg_store_fields(const SqgDev& d, int m, int r, int c, int which0, int lev, double2 z) {
  double2 v = make_double2(2.25 * z.x, 2.25 * z.y);
  if (c == d.NK - 1) v.y = -v.y;
  const double kx = d.kx[c], ly = d.ly[r];
  const long long ld = (long long)d.nm * 8 * d.NK;
  const long long im_off = (long long)d.N * ld;
  const long long base = (long long)r * ld + ((long long)m * 8) * d.NK + c;
  const long long fx = base + (long long)(which0 * 2 + lev) * d.NK, fy = base + (long long)((which0 + 1) * 2 + lev) * d.NK;
  d.spec8[fx] = -kx * v.y; d.spec8[fx + im_off] = kx * v.x;      // i kx (a + i b) = -kx b + i kx a
  d.spec8[fy] = -ly * v.y; d.spec8[fy + im_off] = ly * v.x;
}

// stage: -1 = first prep of a forecast (xs = x); 0..3 = finish RK stage `stage` from the Jacobian in jst, then
// prepare the next stage (do_prep = 0 after the very last stage)
__global__ void __launch_bounds__(128)
sqg_stage_kernel(const SqgDev d, int stage, int do_prep) {
  const long long per = (long long)d.N * d.NK;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * d.nm) return;
  const int m = (int)(idx / per);
  const int rc = (int)(idx - (long long)m * per);
  const int r = rc / d.NK, c = rc - r * d.NK;
  const long long i0 = ((long long)m * 2) * per + rc, i1 = i0 + per;
  double2 s0, s1;                                           // the state the next rhs is evaluated at
  if (stage < 0) {
    s0 = d.x[i0]; s1 = d.x[i1];
    d.xs[i0] = s0; d.xs[i1] = s1;
  } else {
    const double2 x0 = d.x[i0], x1 = d.x[i1], e0 = d.xs[i0], e1 = d.xs[i1];
    // Jacobian spectrum, truncated (_spectrunc, :311-325): the Nyquist column is zero
    double2 j0 = make_double2(0.0, 0.0), j1 = j0;
    const int Nh = d.N / 2;
    if (c < Nh) {
      const long long jb = (((long long)m * 2) * 2 * d.N + r) * Nh + c, lev = (long long)2 * d.N * Nh, im = (long long)d.N * Nh;
      j0 = make_double2(d.jst[jb], d.jst[jb + im]);
      j1 = make_double2(d.jst[jb + lev], d.jst[jb + lev + im]);
    }
    // rhs, :530-545
    double2 t0 = csub(cmul_r(d.inv_tdiab, csub(d.pveq[rc], e0)), j0);
    double2 t1 = csub(cmul_r(d.inv_tdiab, csub(d.pveq[per + rc], e1)), j1);
    if (d.ekman) {
      double2 p0, p1;
      sqg_invert(d, rc, e0, e1, &p0, &p1);
      const double rk = d.r * d.ksqlsq[rc];
      t0 = cadd(t0, cmul_r(rk, p0));
      if (d.symmetric) t1 = cadd(t1, cmul_r(-1.0 * rk, p1));
    }
    const double2 k0 = cmul_r(d.dt, t0), k1 = cmul_r(d.dt, t1);
    // _rk4, :345-364: y = x + (k1 + 2 k2 + 2 k3 + k4) / 6, summed left to right
    if (stage == 0) {
      d.acc[i0] = k0; d.acc[i1] = k1;
      s0 = cadd(x0, cmul_r(0.5, k0)); s1 = cadd(x1, cmul_r(0.5, k1));
    } else if (stage < 3) {
      d.acc[i0] = cadd(d.acc[i0], cmul_r(2.0, k0)); d.acc[i1] = cadd(d.acc[i1], cmul_r(2.0, k1));
      const double w = stage == 1 ? 0.5 : 1.0;
      s0 = cadd(x0, cmul_r(w, k0)); s1 = cadd(x1, cmul_r(w, k1));
    } else {
      const double hd = d.hyperdiff[rc];
      s0 = cmul_r(hd, cadd(x0, cdiv_r(cadd(d.acc[i0], k0), 6.0)));
      s1 = cmul_r(hd, cadd(x1, cdiv_r(cadd(d.acc[i1], k1), 6.0)));
      d.x[i0] = s0; d.x[i1] = s1;
      if (d.st2 != nullptr) {
        const long long ld = (long long)d.nm * 2 * d.NK, im = (long long)d.N * ld;
        const long long o = (long long)r * ld + ((long long)m * 2) * d.NK + c;
        d.st2[o] = s0.x; d.st2[o + im] = s0.y;
        d.st2[o + d.NK] = s1.x; d.st2[o + d.NK + im] = s1.y;
      }
    }
    d.xs[i0] = s0; d.xs[i1] = s1;
  }
  if (!do_prep) return;
  double2 p0, p1;
  sqg_invert(d, rc, s0, s1, &p0, &p1);
  sqg_store_fields(d, m, r, c, 0, 0, p0);
  sqg_store_fields(d, m, r, c, 0, 1, p1);
  sqg_store_fields(d, m, r, c, 2, 0, s0);
  sqg_store_fields(d, m, r, c, 2, 1, s1);
}

// state [nm][2][N][NK] complex -> the [2 (re, im)][N][nm][2][NK] layout of the trajectory transform
__global__ void __launch_bounds__(128)
sqg_pack_kernel(const double2* __restrict__ x, double* __restrict__ st2, int N, int NK, int nm) {
  const long long per = (long long)N * NK;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * nm) return;
  const int m = (int)(idx / per);
  const int rc = (int)(idx - (long long)m * per);
  const int r = rc / NK, c = rc - r * NK;
  const long long ld = (long long)nm * 2 * NK, im = (long long)N * ld;
  const long long o = (long long)r * ld + ((long long)m * 2) * NK + c;
  const double2 s0 = x[((long long)m * 2) * per + rc], s1 = x[((long long)m * 2 + 1) * per + rc];
  st2[o] = s0.x; st2[o + im] = s0.y;
  st2[o + NK] = s1.x; st2[o + NK + im] = s1.y;
}

// F [nm][8][M][M] (field = which * 2 + level) -> J [nm][2][M][M] = psi_x q_y - psi_y q_x
__global__ void __launch_bounds__(256)
sqg_jacobian_kernel(const double* __restrict__ F, double* __restrict__ J, long long mm, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long ml = idx / mm, p = idx - ml * mm;
  const long long m = ml >> 1, lev = ml & 1;
  const double* f = F + (m * 8 + lev) * mm + p;
  J[idx] = f[0] * f[6 * mm] - f[2 * mm] * f[4 * mm];
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
struct SqgState {
  int N = 0, NK = 0, M = 0;
  SqgDev dev{};
  std::vector<void*> owned;            // tables and constant matrices
  double *Wb = nullptr, *Cc = nullptr, *Sc = nullptr, *Df = nullptr, *We1 = nullptr, *We2 = nullptr;
  double *Wt = nullptr, *Ct = nullptr, *St = nullptr;
  double *Dg = nullptr, *Vg1 = nullptr, *Vg2 = nullptr;     // rfft2 on the N grid (gridded states in)
  // per-ensemble workspaces (grow with the member count)
  int cap = 0;
  double *x = nullptr, *xs = nullptr, *acc = nullptr, *spec8 = nullptr, *yst = nullptr, *F = nullptr, *J = nullptr,
         *G = nullptr, *jst = nullptr, *st2 = nullptr, *yt = nullptr, *gg = nullptr;
};

static cudaError_t upload(SqgState* s, const std::vector<double>& v, double** out) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, sizeof(double) * (v.empty() ? 1 : v.size()));
  if (e != cudaSuccess) return e;
  s->owned.push_back(p);
  *out = static_cast<double*>(p);
  return cudaMemcpy(p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
}

static void free_work(SqgState* s) {
  double** w[] = {&s->x, &s->xs, &s->acc, &s->spec8, &s->yst, &s->F, &s->J, &s->G, &s->jst, &s->st2, &s->yt, &s->gg};
  for (double** p : w) { if (*p) cudaFree(*p); *p = nullptr; }
  s->cap = 0;
}

void sqg_destroy(SqgState* s) {
  if (s == nullptr) return;
  free_work(s);
  for (void* p : s->owned) cudaFree(p);
  delete s;
}

// cos / sin of 2 pi a b / n with the argument reduced exactly (a b mod n in integers)
static inline void unit(long long a, long long b, long long n, double* c, double* s) {
  long long q = (a * b) % n;
  if (q < 0) q += n;
  const double ang = 2.0 * M_PI * (double)q / (double)n;
  *c = cos(ang); *s = sin(ang);
}

cudaError_t sqg_setup(SqgState** out, const SqgTables& t) {
  SqgState* s = new SqgState();
  const int N = t.n, Nh = N / 2, NK = Nh + 1, M = 3 * N / 2;
  s->N = N; s->NK = NK; s->M = M;
  cudaError_t e = cudaSuccess;
  auto up = [&](const double* src, size_t n, const double** dst) {
    if (e != cudaSuccess) return;
    std::vector<double> v(src, src + n);
    double* p = nullptr;
    e = upload(s, v, &p);
    *dst = p;
  };
  const size_t per = (size_t)N * NK;
  up(t.kx, NK, &s->dev.kx); up(t.ly, N, &s->dev.ly);
  up(t.hovermu, per, &s->dev.hovermu); up(t.tanhmu, per, &s->dev.tanhmu); up(t.sinhmu, per, &s->dev.sinhmu);
  up(t.ksqlsq, per, &s->dev.ksqlsq); up(t.hyperdiff, per, &s->dev.hyperdiff);
  const double* pveq = nullptr;
  up(t.pvspec_eq, 4 * per, &pveq);
  s->dev.pveq = reinterpret_cast<const double2*>(pveq);
  s->dev.N = N; s->dev.NK = NK; s->dev.M = M;
  s->dev.ekman = t.ekman; s->dev.symmetric = t.symmetric;
  s->dev.dt = t.delta_t; s->dev.inv_tdiab = t.inv_tdiab; s->dev.r = t.r;
  auto wave = [&](int r) { return r < Nh ? r : r - N; };        // signed wavenumber of kept row r (both grids)
  // gemm b: [Yr; Yi] = [[Wr, -Wi], [Wi, Wr]] [Ar; Ai], W[y][r] = exp(+2 pi i y l_r / M)
  {
    std::vector<double> W((size_t)2 * M * 2 * N);
    for (int y = 0; y < M; ++y)
      for (int r = 0; r < N; ++r) {
        double c, sn; unit(y, wave(r), M, &c, &sn);
        W[(size_t)y * 2 * N + r] = c;            W[(size_t)y * 2 * N + N + r] = -sn;
        W[(size_t)(M + y) * 2 * N + r] = sn;     W[(size_t)(M + y) * 2 * N + N + r] = c;
      }
    if (e == cudaSuccess) e = upload(s, W, &s->Wb);
  }
  // gemm c: f[y][x] = sum_c Yr[y][c] Cc[c][x] + Yi[y][c] Sc[c][x]; weights 1 (c = 0) / 2, 1 / M^2, Sc = -w sin
  {
    std::vector<double> C((size_t)NK * M), S((size_t)NK * M);
    const double nrm = 1.0 / ((double)M * (double)M);
    for (int c = 0; c < NK; ++c)
      for (int x = 0; x < M; ++x) {
        double cs, sn; unit(c, x, M, &cs, &sn);
        const double w = (c == 0 ? 1.0 : 2.0) * nrm;
        C[(size_t)c * M + x] = w * cs; S[(size_t)c * M + x] = -w * sn;
      }
    if (e == cudaSuccess) e = upload(s, C, &s->Cc);
    if (e == cudaSuccess) e = upload(s, S, &s->Sc);
  }
  // gemm d: G[y][kc] = sum_x J[y][x] cos, G[y][Nh + kc] = -sum_x J[y][x] sin  (exp(-2 pi i kc x / M))
  {
    std::vector<double> D((size_t)M * N);
    for (int x = 0; x < M; ++x)
      for (int kc = 0; kc < Nh; ++kc) {
        double cs, sn; unit(kc, x, M, &cs, &sn);
        D[(size_t)x * N + kc] = cs; D[(size_t)x * N + Nh + kc] = -sn;
      }
    if (e == cudaSuccess) e = upload(s, D, &s->Df);
  }
  // gemm e: [Jr; Ji] = [Vr; Vi] Gr + [-Vi; Vr] Gi, V[r][y] = exp(-2 pi i l_r y / M)
  {
    std::vector<double> E1((size_t)2 * N * M), E2((size_t)2 * N * M);
    for (int r = 0; r < N; ++r)
      for (int y = 0; y < M; ++y) {
        double cs, sn; unit(wave(r), y, M, &cs, &sn);
        E1[(size_t)r * M + y] = cs;  E1[(size_t)(N + r) * M + y] = -sn;
        E2[(size_t)r * M + y] = sn;  E2[(size_t)(N + r) * M + y] = cs;
      }
    if (e == cudaSuccess) e = upload(s, E1, &s->We1);
    if (e == cudaSuccess) e = upload(s, E2, &s->We2);
  }
  // trajectory: irfft2 on the N grid (the Nyquist column counts once, its imaginary part is ignored: sin(pi x) = 0)
  {
    std::vector<double> W((size_t)2 * N * 2 * N), C((size_t)NK * N), S((size_t)NK * N);
    for (int y = 0; y < N; ++y)
      for (int r = 0; r < N; ++r) {
        double c, sn; unit(y, wave(r), N, &c, &sn);
        W[(size_t)y * 2 * N + r] = c;            W[(size_t)y * 2 * N + N + r] = -sn;
        W[(size_t)(N + y) * 2 * N + r] = sn;     W[(size_t)(N + y) * 2 * N + N + r] = c;
      }
    const double nrm = 1.0 / ((double)N * (double)N);
    for (int c = 0; c < NK; ++c)
      for (int x = 0; x < N; ++x) {
        double cs, sn; unit(c, x, N, &cs, &sn);
        const double w = ((c == 0 || c == Nh) ? 1.0 : 2.0) * nrm;
        C[(size_t)c * N + x] = w * cs; S[(size_t)c * N + x] = (c == 0 || c == Nh) ? 0.0 : -w * sn;
      }
    if (e == cudaSuccess) e = upload(s, W, &s->Wt);
    if (e == cudaSuccess) e = upload(s, C, &s->Ct);
    if (e == cudaSuccess) e = upload(s, S, &s->St);
  }
  // rfft2 on the N grid: along x  G[y][c] = sum_x f[y][x] exp(-2 pi i c x / N), c = 0 .. N/2 (real | imaginary
  // halves of a row), then along y  X[r][c] = sum_y exp(-2 pi i l_r y / N) G[y][c]
  {
    std::vector<double> D((size_t)N * 2 * NK), V1((size_t)2 * N * N), V2((size_t)2 * N * N);
    for (int x = 0; x < N; ++x)
      for (int c = 0; c < NK; ++c) {
        double cs, sn; unit(c, x, N, &cs, &sn);
        D[(size_t)x * 2 * NK + c] = cs; D[(size_t)x * 2 * NK + NK + c] = -sn;
      }
    for (int r = 0; r < N; ++r)
      for (int y = 0; y < N; ++y) {
        double cs, sn; unit(wave(r), y, N, &cs, &sn);
        V1[(size_t)r * N + y] = cs;  V1[(size_t)(N + r) * N + y] = -sn;
        V2[(size_t)r * N + y] = sn;  V2[(size_t)(N + r) * N + y] = cs;
      }
    if (e == cudaSuccess) e = upload(s, D, &s->Dg);
    if (e == cudaSuccess) e = upload(s, V1, &s->Vg1);
    if (e == cudaSuccess) e = upload(s, V2, &s->Vg2);
  }
  if (e != cudaSuccess) { sqg_destroy(s); return e; }
  *out = s;
  return cudaSuccess;
}

static cudaError_t reserve_work(SqgState* s, int nm) {
  if (nm <= s->cap) return cudaSuccess;
  free_work(s);
  const size_t N = s->N, NK = s->NK, M = s->M, per = 2 * N * NK;
  struct { double** p; size_t n; } w[] = {
      {&s->x, 2 * per * nm}, {&s->xs, 2 * per * nm}, {&s->acc, 2 * per * nm},
      {&s->spec8, 2 * N * nm * 8 * NK}, {&s->yst, 2 * M * nm * 8 * NK}, {&s->F, (size_t)nm * 8 * M * M},
      {&s->J, (size_t)nm * 2 * M * M}, {&s->G, (size_t)nm * 2 * M * N}, {&s->jst, (size_t)nm * 2 * 2 * N * (N / 2)},
      {&s->st2, 2 * N * nm * 2 * NK}, {&s->yt, 2 * N * nm * 2 * NK}, {&s->gg, (size_t)nm * 2 * N * 2 * NK}};
  for (auto& b : w) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(b.p), sizeof(double) * b.n);
    if (e != cudaSuccess) { free_work(s); return e; }
  }
  s->cap = nm;
  return cudaSuccess;
}

// spec_in / spec_out: [members][2][N][N/2+1] complex128 (device; may alias); grid_in / grid_out: [members][2][N][N]
// real (the state enters as rfft2(grid_in) when spec_in is null; grid_out = irfft2 of the final state); spec_out and
// grid_out are both optional; traj: nullable [n_steps][members][2][N][N]
cudaError_t sqg_forecast(SqgState* s, const double* spec_in, const double* grid_in, double* spec_out, double* grid_out,
                         double* traj, int nm, int n_steps, cudaStream_t st, int* n_launches) {
  cudaError_t e = reserve_work(s, nm);
  if (e != cudaSuccess) return e;
  const int N = s->N, Nh = N / 2, NK = s->NK, M = s->M;
  const size_t state_bytes = sizeof(double) * 2 * 2 * N * NK * (size_t)nm;
  int nl = 0;
  if (spec_in != nullptr) {
    e = cudaMemcpyAsync(s->x, spec_in, state_bytes, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  } else {
    // rfft2: rows (member, level, y) x [N] times [N x 2 NK], then per (member, level) the y transform, the real and
    // the imaginary rows written straight into the interleaved complex state
    e = gemm(st, 1, nm * 2 * N, 2 * NK, N, {grid_in, N, 0}, {s->Dg, 2 * NK, 0}, s->gg, 2 * NK, 0);
    if (e != cudaSuccess) return e;
    const long long gb = (long long)N * 2 * NK, xb = (long long)N * NK * 2;
    for (int ri = 0; ri < 2; ++ri) {
      GemmOp a2{s->Vg2 + (size_t)ri * N * N, N, 0}, b2{s->gg + NK, 2 * NK, gb};
      e = gemm(st, nm * 2, N, NK, N, {s->Vg1 + (size_t)ri * N * N, N, 0}, {s->gg, 2 * NK, gb}, s->x + ri, 2 * NK, xb, &a2, &b2, 2);
      if (e != cudaSuccess) return e;
    }
    nl += 3;
  }
  SqgDev d = s->dev;
  d.nm = nm;
  d.x = reinterpret_cast<double2*>(s->x); d.xs = reinterpret_cast<double2*>(s->xs); d.acc = reinterpret_cast<double2*>(s->acc);
  const bool want_grid = traj != nullptr || (grid_out != nullptr && n_steps > 0);
  d.spec8 = s->spec8; d.jst = s->jst; d.st2 = want_grid ? s->st2 : nullptr;
  const long long pts = (long long)nm * N * NK;
  const unsigned sgrid = (unsigned)((pts + 127) / 128);
  const long long ld8 = (long long)nm * 8 * NK, ld2 = (long long)nm * 2 * NK;
  const long long mm = (long long)M * M, jtot = (long long)nm * 2 * mm;
  if (n_steps > 0) {
    sqg_stage_kernel<<<sgrid, 128, 0, st>>>(d, -1, 1); ++nl;
  }
  for (int step = 0; step < n_steps; ++step) {
    for (int stage = 0; stage < 4; ++stage) {
      // b: along y
      e = gemm(st, 1, 2 * M, (int)ld8, 2 * N, {s->Wb, 2 * N, 0}, {s->spec8, ld8, 0}, s->yst, ld8, 0);
      if (e != cudaSuccess) return e;
      // c: along x, batch = (member, field)
      {
        GemmOp a2{s->yst + (long long)M * ld8, ld8, NK}, b2{s->Sc, M, 0};
        e = gemm(st, nm * 8, M, M, NK, {s->yst, ld8, NK}, {s->Cc, M, 0}, s->F, M, mm, &a2, &b2);
        if (e != cudaSuccess) return e;
      }
      sqg_jacobian_kernel<<<(unsigned)((jtot + 255) / 256), 256, 0, st>>>(s->F, s->J, mm, jtot);
      // d: forward along x, all (member, level, y) rows at once
      e = gemm(st, 1, nm * 2 * M, N, M, {s->J, M, 0}, {s->Df, N, 0}, s->G, N, 0);
      if (e != cudaSuccess) return e;
      // e: forward along y, batch = (member, level)
      {
        GemmOp a2{s->We2, M, 0}, b2{s->G + Nh, N, (long long)M * N};
        e = gemm(st, nm * 2, 2 * N, Nh, M, {s->We1, M, 0}, {s->G, N, (long long)M * N}, s->jst, Nh, (long long)2 * N * Nh, &a2, &b2);
        if (e != cudaSuccess) return e;
      }
      const bool last = step == n_steps - 1 && stage == 3;
      sqg_stage_kernel<<<sgrid, 128, 0, st>>>(d, stage, last ? 0 : 1);
      nl += 6;
    }
    const bool final_grid = grid_out != nullptr && step == n_steps - 1;
    if (traj != nullptr || final_grid) {
      double* out = traj != nullptr ? traj + (size_t)step * nm * 2 * N * N : grid_out;
      e = gemm(st, 1, 2 * N, (int)ld2, 2 * N, {s->Wt, 2 * N, 0}, {s->st2, ld2, 0}, s->yt, ld2, 0);
      if (e != cudaSuccess) return e;
      GemmOp a2{s->yt + (long long)N * ld2, ld2, NK}, b2{s->St, N, 0};
      e = gemm(st, nm * 2, N, N, NK, {s->yt, ld2, NK}, {s->Ct, N, 0}, out, N, (long long)N * N, &a2, &b2);
      if (e != cudaSuccess) return e;
      nl += 2;
      if (traj != nullptr && final_grid) {
        e = cudaMemcpyAsync(grid_out, out, sizeof(double) * (size_t)nm * 2 * N * N, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
      }
    }
  }
  if (grid_out != nullptr && n_steps == 0) {
    // no step: the gridded state is the projection irfft2(rfft2(.)) of the input
    sqg_pack_kernel<<<(unsigned)(((long long)nm * N * NK + 127) / 128), 128, 0, st>>>(
        reinterpret_cast<const double2*>(s->x), s->st2, N, NK, nm);
    e = gemm(st, 1, 2 * N, (int)ld2, 2 * N, {s->Wt, 2 * N, 0}, {s->st2, ld2, 0}, s->yt, ld2, 0);
    if (e != cudaSuccess) return e;
    GemmOp a2{s->yt + (long long)N * ld2, ld2, NK}, b2{s->St, N, 0};
    e = gemm(st, nm * 2, N, N, NK, {s->yt, ld2, NK}, {s->Ct, N, 0}, grid_out, N, (long long)N * N, &a2, &b2);
    if (e != cudaSuccess) return e;
    nl += 3;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (spec_out != nullptr) e = cudaMemcpyAsync(spec_out, s->x, state_bytes, cudaMemcpyDeviceToDevice, st);
  if (n_launches) *n_launches = nl;
  return e;
}

}  // namespace dab
