// Device-side pieces of K4 shared by the standalone eigensolver kernels (etkf_eig.cu) and the
// batched small-problem kernel (etkf_batched.cu).
#pragma once
#include "common.cuh"

namespace dab {

// -DDAB_EIG_TIMING: warp 0 accumulates clock64() per section into lam_out[0..5] (debug builds only).
#ifdef DAB_EIG_TIMING
#define DAB_TICK(acc) do { const long long now_ = clock64(); (acc) += now_ - tick_; tick_ = now_; } while (0)
#else
#define DAB_TICK(acc) do { } while (0)
#endif

template <int LPP>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// After convergence: column c of G has norm lambda_c and direction v_c.  Builds T in Tout.
// G: [ncol][LD] column-major in shared memory with ROWS >= 8*ceil(k/8) zero-padded rows;
// aux: 6*ncol doubles of shared memory.
template <int LD, int ROWS>
__device__ __forceinline__ void build_transform(double* G, double* aux, int ncol, const double* __restrict__ stats,
                                                int k, double* __restrict__ Tout, double* __restrict__ lam_out,
                                                int* __restrict__ info_out, int sweeps) {
  double* lam = aux;                           // ncol
  double* dw = lam + ncol;                     // ncol  sqrt((k-1)/lambda), 0 beyond k
  double* da = dw + ncol;                      // ncol  1/lambda (0 below the pinv cutoff)
  double* wa = da + ncol;                      // ncol
  double* proj = wa + ncol;                    // ncol  v_c . g
  double* cmean = proj + ncol;                 // ncol
  __shared__ double s_red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const double inv_k = 1.0 / (double)k;
  const int k8 = (k + 7) >> 3;
  // eigenvalues = column norms; normalise columns to eigenvectors; proj = V^T g
  const double* gvec = stats + k * k;
  for (int c = warp; c < k; c += nwarp) {
    double* gc = G + c * LD;
    double ss = 0.0;
    for (int r = lane; r < ROWS; r += 32) ss = fma(gc[r], gc[r], ss);
    ss = group_sum<32>(ss);
    const double l = sqrt(ss);
    const double il = l > 0.0 ? 1.0 / l : 0.0;
    double pg = 0.0;
    for (int r = lane; r < ROWS; r += 32) {
      const double v = gc[r] * il;
      gc[r] = v;
      if (r < k) pg = fma(v, gvec[r], pg);
    }
    pg = group_sum<32>(pg);
    if (lane == 0) { lam[c] = l; proj[c] = pg; }
  }
  __syncthreads();
  // pinv cutoff: 1e-15 * largest eigenvalue (rtol semantics of jnp.linalg.pinv)
  double lmax = 0.0;
  for (int c = tid; c < k; c += blockDim.x) lmax = fmax(lmax, lam[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) s_red[warp] = lmax;
  __syncthreads();
  lmax = 0.0;
  for (int w = 0; w < nwarp; ++w) lmax = fmax(lmax, s_red[w]);
  for (int c = tid; c < ncol; c += blockDim.x) {
    double inv = 0.0, l = 0.0;
    if (c < k) {
      l = lam[c];
      inv = l > 1e-15 * lmax ? 1.0 / l : 0.0;
      if (lam_out) lam_out[c] = l;
    }
    da[c] = inv;
    dw[c] = sqrt((double)(k - 1) * inv);
  }
  __syncthreads();
  // wa = V diag(da) V^T g
  for (int r = tid; r < k; r += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < k; ++c) s = fma(G[c * LD + r] * da[c], proj[c], s);
    wa[r] = s;
  }
  // column means of Wa without forming it: cmean_j = (1/k) sum_c V[j][c] dw_c (1^T v_c)
  for (int c = warp; c < k; c += nwarp) {
    double u = 0.0;
    for (int r = lane; r < ROWS; r += 32) u += G[c * LD + r];
    u = group_sum<32>(u);
    if (lane == 0) lam[c] = u * dw[c];             // lam[] is free now (copied to lam_out above)
  }
  __syncthreads();
  for (int r = tid; r < k; r += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < k; ++c) s = fma(G[c * LD + r], lam[c], s);
    cmean[r] = s * inv_k;
  }
  double wsum = 0.0;
  for (int r = 0; r < k; ++r) wsum += wa[r];       // every thread, same order: identical value
  const double wmean = wsum * inv_k;
  __syncthreads();
  // Wa = (V diag(dw)) V^T on the tensor core, epilogue writes
  // T[m][j] = 1/k + (wa[m] - mean(wa)) + (Wa[m][j] - colmean_j).  Strip `st` = rows 8*st..8*st+7.
  {
    const int fm = lane >> 2, fr = lane & 3;
    for (int st = warp; st < k8; st += nwarp) {
      const int row = 8 * st + fm;
      const double base = row < k ? inv_k + (wa[row] - wmean) : 0.0;
      for (int tj = 0; tj < k8; ++tj) {
        double c0 = 0.0, c1 = 0.0;
        for (int ks = 0; ks < 2 * k8; ++ks) {
          const int c = 4 * ks + fr;
          const double a = G[c * LD + 8 * st + fm] * dw[c];
          const double b = G[c * LD + 8 * tj + fm];
          dmma884(c0, c1, a, b);
        }
        const int col = 8 * tj + 2 * fr;
        if (row < k) {
          if (col < k) Tout[row * k + col] = base + (c0 - cmean[col]);
          if (col + 1 < k) Tout[row * k + col + 1] = base + (c1 - cmean[col + 1]);
        }
      }
    }
  }
  if (info_out && tid == 0) *info_out = sweeps;
}

// stats: [k*k] C row-major, then [k] g.  Tout: [k][k] row-major, T[m][j].
// lam_out (nullable): k eigenvalues of A (unsorted); info_out (nullable): sweeps used.
// `sm`: eig_small_smem_doubles(k) doubles of shared memory.  stats / Tout may live in global or
// shared memory.  Must be called by every thread of the CTA (it synchronises the CTA).
template <int LPP>
__device__ __forceinline__ void eig_small_body(const double* stats, int k, double rho, int empty_R,
                                               double* Tout, double* lam_out, int* info_out, double* sm) {
  constexpr int RPL = 8;                       // rows per lane
  constexpr int ROWS = RPL * LPP;              // padded row count (>= k)
  constexpr int LD = ROWS + 8;                 // column stride: keeps the 4 pairs of a warp 2-way at worst
  constexpr int PPW = 32 / LPP;                // pairs per warp
  const int kj = (k + 1) & ~1;                 // even number of Jacobi columns (last may be a zero column)
  const int k8 = (k + 7) >> 3;
  const int ncol = kj > 8 * k8 ? kj : 8 * k8;  // columns kept (zero beyond k) so DMMA tiles need no guards
  double* G = sm;                              // ncol * LD, column-major
  double* aux = G + ncol * LD;                 // 6 * ncol

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv_k = 1.0 / (double)k;

  if (empty_R) {                               // reference zero branch: T = U
    for (int e = tid; e < k * k; e += blockDim.x) Tout[e] = inv_k;
    if (info_out && tid == 0) *info_out = 0;
    __syncthreads();
    return;
  }

  const double shift = (double)(k - 1) / rho;
  for (int e = tid; e < ncol * LD; e += blockDim.x) {
    const int c = e / LD, r = e - c * LD;
    double v = 0.0;
    if (c < k && r < k) v = stats[r * k + c] + (r == c ? shift : 0.0);
    G[e] = v;
  }
  __syncthreads();

  const int npairs = kj / 2;
  const int m = kj - 1;                        // round-robin modulus
  const int sub = lane % LPP;
  const int pr = warp * PPW + lane / LPP;      // pair slot of this lane group
  const bool active = pr < npairs;
  const double tol = 1.1e-16 * sqrt((double)k) * 4.0;
  const double tol2 = tol * tol;
  const double quad2 = 1e-16;                  // (1e-8)^2
  int sweeps = 0;
  for (; sweeps < 40; ++sweeps) {
    int rotated = 0, big = 0;
    for (int rnd = 0; rnd < m; ++rnd) {
      {
        // every lane runs the loads and shuffles (full-mask shuffles need the whole warp);
        // lane groups without a pair work on columns (0,1) and discard the result.
        int p = 0, q = 1;
        if (active) {
          if (pr == 0) { p = m; q = rnd; }
          else {
            p = rnd + pr; if (p >= m) p -= m;
            q = rnd - pr; if (q < 0) q += m;
          }
          if (p > q) { const int t_ = p; p = q; q = t_; }
        }
        double* gp = G + p * LD + sub;
        double* gq = G + q * LD + sub;
        double xp[RPL], xq[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) { xp[i] = gp[LPP * i]; xq[i] = gq[LPP * i]; }
        double al0 = xp[0] * xp[0], be0 = xq[0] * xq[0], ga0 = xp[0] * xq[0];
        double al1 = xp[1] * xp[1], be1 = xq[1] * xq[1], ga1 = xp[1] * xq[1];
#pragma unroll
        for (int i = 2; i < RPL; i += 2) {
          al0 = fma(xp[i], xp[i], al0); be0 = fma(xq[i], xq[i], be0); ga0 = fma(xp[i], xq[i], ga0);
          al1 = fma(xp[i + 1], xp[i + 1], al1); be1 = fma(xq[i + 1], xq[i + 1], be1);
          ga1 = fma(xp[i + 1], xq[i + 1], ga1);
        }
        const double al = group_sum<LPP>(al0 + al1);
        const double be = group_sum<LPP>(be0 + be1);
        const double ga = group_sum<LPP>(ga0 + ga1);
        const double g2 = ga * ga, ab = al * be;
        if (active && g2 > tol2 * ab) {        // uniform within the lane group
          rotated = 1;
          if (g2 > quad2 * ab) big = 1;
          const double tau = be - al;
          const double r = rsqrt(fma(tau, tau, 4.0 * g2));
          const double x = fma(0.5 * fabs(tau), r, 0.5);
          const double ic = rsqrt(x);
          const double c_ = x * ic;
          const double s_ = copysign(ga * r * ic, tau >= 0.0 ? ga : -ga);
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            gp[LPP * i] = fma(c_, xp[i], -s_ * xq[i]);
            gq[LPP * i] = fma(s_, xp[i], c_ * xq[i]);
          }
        }
      }
      __syncthreads();
    }
    const int any_rot = __syncthreads_or(rotated);
    const int any_big = __syncthreads_or(big);
    if (!any_rot) break;
    if (!any_big) { ++sweeps; break; }
  }

  build_transform<LD, ROWS>(G, aux, ncol, stats, k, Tout, lam_out, info_out, sweeps);
}


// ---- warp-local block Jacobi (32 < k <= 128) ------------------------------------------------
// Columns are grouped in blocks of 4; a warp keeps a PAIR of blocks (8 columns, 8 lanes x RPL rows
// per column pair) in registers for a whole block-round and rotates the 16 cross pairs in four
// sub-rounds, passing the second block's columns from lane group to lane group with shuffles.
// Shared memory is touched, and the CTA synchronised, once per block-round (15 per sweep at
// k = 64) instead of once per rotation round (63); the three rounds of within-block pairs go
// through shared memory.  Column norms ride along with the columns (updated by the rotation,
// recomputed at every load), so a sub-round needs one inner product instead of three.
// `ncol` (multiple of 8, >= k) columns are kept, zero beyond k; the CTA must have 32 * ncol / 8
// threads.  `sm`: eig_block_smem_doubles(ncol, ROWS) doubles of shared memory.
template <int RPL>
__device__ __forceinline__ void eig_block_body(const double* __restrict__ stats, int k, int ncol, double rho,
                                               double* __restrict__ Tout, double* __restrict__ lam_out,
                                               int* __restrict__ info_out, int skew, double* sm) {
  constexpr int LPP = 8;
  constexpr int ROWS = RPL * LPP;
  constexpr int LD = ROWS + 8;
  const int nb = ncol >> 2;
  double* G = sm;
  double* aux = G + ncol * LD;
  double* nrm = aux + 6 * ncol;                // squared column norms, carried with the columns
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, grp = lane >> 3;

  const double shift = (double)(k - 1) / rho;
  for (int e = tid; e < ncol * LD; e += blockDim.x) {
    const int c = e / LD, r = e - c * LD;
    double v = 0.0;
    if (c < k && r < k) v = stats[r * k + c] + (r == c ? shift : 0.0);
    G[e] = v;
  }
  __syncthreads();

  const double tol = 1.1e-16 * sqrt((double)k) * 4.0;
  const double tol2 = tol * tol;
  const double quad2 = 1e-16;
  const int m = nb - 1;
  int sweeps = 0;
#ifdef DAB_EIG_TIMING
  long long tA = 0, tLoad = 0, tSub = 0, tStore = 0, tFlags = 0, tick_ = clock64();
#endif
  for (; sweeps < 40; ++sweeps) {
    int rotated = 0, big = 0;
    DAB_TICK(tFlags);
    // ---- phase A: the 6 pairs inside every block, 3 rounds through shared memory
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      const int gid = warp * 4 + grp;           // 2 lane groups per block
      const int blk = gid >> 1, wh = gid & 1;
      int p, q;
      if (t == 0) { p = 2 * wh; q = 2 * wh + 1; }
      else if (t == 1) { p = wh; q = wh + 2; }
      else { p = wh; q = 3 - wh; }
      double* gp = G + (4 * blk + p) * LD + sub;
      double* gq = G + (4 * blk + q) * LD + sub;
      double xp[RPL], xq[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) { xp[i] = gp[LPP * i]; xq[i] = gq[LPP * i]; }
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        al = fma(xp[i], xp[i], al); be = fma(xq[i], xq[i], be); ga = fma(xp[i], xq[i], ga);
      }
      al = group_sum<LPP>(al); be = group_sum<LPP>(be); ga = group_sum<LPP>(ga);
      const double g2 = ga * ga, ab = al * be;
      if (g2 > tol2 * ab) {
        rotated = 1;
        if (g2 > quad2 * ab) big = 1;
        const double tau = be - al;
        const double q_ = fma(tau, tau, 4.0 * g2);
        const double r = rsqrt(q_);
        const double x = fma(0.5 * fabs(tau), r, 0.5);
        const double ic = rsqrt(x);
        const double c_ = x * ic;
        const double s_ = copysign(ga * r * ic, tau >= 0.0 ? ga : -ga);
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          gp[LPP * i] = fma(c_, xp[i], -s_ * xq[i]);
          gq[LPP * i] = fma(s_, xp[i], c_ * xq[i]);
        }
        // the rotation diagonalises the 2x2 Gram: norms' = (al + be -/+ sgn(tau) h) / 2, h = sqrt(q)
        const double hh = copysign(q_ * r, tau);
        const double sum = al + be;
        al = 0.5 * (sum - hh); be = 0.5 * (sum + hh);
      }
      if (t == 2 && sub == 0) { nrm[4 * blk + p] = al; nrm[4 * blk + q] = be; }
      __syncthreads();
    }
    DAB_TICK(tA);
    // ---- phase B: block round-robin, 16 cross pairs per block pair in registers
#pragma unroll 1
    for (int R = 0; R < m; ++R) {
      int I, J;
      if (warp == 0) { I = m; J = R; }
      else {
        I = R + warp; if (I >= m) I -= m;
        J = R - warp; if (J < 0) J += m;
      }
      if (skew > 0 && (warp & 4)) {              // de-phase the two warps that share an SMSP
        const long long t_end = clock64() + skew;
        while (clock64() < t_end) { }
      }
      double* ga_ptr = G + (4 * I + grp) * LD + sub;
      double xa[RPL], xb[RPL];
      double na = nrm[4 * I + grp], nbn = nrm[4 * J + grp];
      {
        const double* gb_ptr = G + (4 * J + grp) * LD + sub;
#pragma unroll
        for (int i = 0; i < RPL; ++i) { xa[i] = ga_ptr[LPP * i]; xb[i] = gb_ptr[LPP * i]; }
      }
#ifdef DAB_EIG_TIMING
      if (xa[0] + xb[0] + na == -1.2345e300) tLoad -= 1;   // make the loads complete before the tick
#endif
      DAB_TICK(tLoad);
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        double g0 = xa[0] * xb[0], g1 = xa[1] * xb[1];
#pragma unroll
        for (int i = 2; i < RPL; i += 2) { g0 = fma(xa[i], xb[i], g0); g1 = fma(xa[i + 1], xb[i + 1], g1); }
        const double ga = group_sum<LPP>(g0 + g1);
        const double g2 = ga * ga, ab = na * nbn;
        const bool rot = g2 > tol2 * ab;
        const int src = (lane + 8) & 31;         // the b columns move to the previous lane group
        if (__any_sync(0xffffffffu, rot)) {      // warp-uniform: lets the shuffles overlap the rotation
          double c_ = 1.0, s_ = 0.0;
          if (rot) {
            rotated = 1;
            if (g2 > quad2 * ab) big = 1;
            const double tau = nbn - na;
            const double q_ = fma(tau, tau, 4.0 * g2);
            const double r = rsqrt(q_);
            const double x = fma(0.5 * fabs(tau), r, 0.5);
            const double ic = rsqrt(x);
            c_ = x * ic;
            s_ = copysign(ga * r * ic, tau >= 0.0 ? ga : -ga);
            // norms' = (na + nb -/+ sgn(tau) h) / 2 with h = sqrt(q) = q r: off the c/s chain
            const double hh = copysign(q_ * r, tau);
            const double sum = na + nbn;
            na = 0.5 * (sum - hh); nbn = 0.5 * (sum + hh);
          }
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const double a_ = xa[i], b_ = xb[i];
            xa[i] = fma(c_, a_, -s_ * b_);
            const double nb_ = fma(s_, a_, c_ * b_);
            xb[i] = (sr < 3) ? __shfl_sync(0xffffffffu, nb_, src) : nb_;
          }
          if (sr < 3) nbn = __shfl_sync(0xffffffffu, nbn, src);
        } else if (sr < 3) {
#pragma unroll
          for (int i = 0; i < RPL; ++i) xb[i] = __shfl_sync(0xffffffffu, xb[i], src);
          nbn = __shfl_sync(0xffffffffu, nbn, src);
        }
      }
#ifdef DAB_EIG_TIMING
      if (xa[0] + xb[0] == -1.2345e300) tSub -= 1;
#endif
      DAB_TICK(tSub);
      double* gb_out = G + (4 * J + ((grp + 3) & 3)) * LD + sub;
#pragma unroll
      for (int i = 0; i < RPL; ++i) { ga_ptr[LPP * i] = xa[i]; gb_out[LPP * i] = xb[i]; }
      if (sub == 0) { nrm[4 * I + grp] = na; nrm[4 * J + ((grp + 3) & 3)] = nbn; }
      __syncthreads();
      DAB_TICK(tStore);
    }
    const int any_rot = __syncthreads_or(rotated);
    const int any_big = __syncthreads_or(big);
    if (!any_rot) break;
    if (!any_big) { ++sweeps; break; }
  }
  DAB_TICK(tFlags);
  build_transform<LD, ROWS>(G, aux, ncol, stats, k, Tout, lam_out, info_out, sweeps);
#ifdef DAB_EIG_TIMING
  long long tBuild = 0;
  DAB_TICK(tBuild);
  if (tid == 0 && lam_out) {
    lam_out[0] = (double)tA; lam_out[1] = (double)tLoad; lam_out[2] = (double)tSub;
    lam_out[3] = (double)tStore; lam_out[4] = (double)tFlags; lam_out[5] = (double)tBuild;
  }
#endif
}


__host__ __device__ inline int eig_small_lpp(int k) { return k <= 8 ? 1 : (k <= 16 ? 2 : (k <= 32 ? 4 : (k <= 64 ? 8 : 16))); }

__host__ __device__ inline int eig_small_smem_doubles(int k) {
  const int lpp = eig_small_lpp(k);
  const int ld = 8 * lpp + 8;
  const int kj = (k + 1) & ~1, k8 = (k + 7) >> 3;
  const int ncol = kj > 8 * k8 ? kj : 8 * k8;
  return ncol * ld + 6 * ncol + 8;
}

}  // namespace dab
