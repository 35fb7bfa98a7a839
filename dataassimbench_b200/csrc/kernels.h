// Internal launch interface between the engine (engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace dab {

constexpr int kMaxRanks = 8;

// Optional second output of the forecast kernels: the observed columns of the forecast, compact and in the cycler's
// location order, Yb[k][ldy] -- what the contraction of the NEXT analysis reads (etkf_contract_dense.cu) instead of
// gathering from the state.  Stationary networks only: the rows whose column lies in a tile's stored range [p, q)
// are a contiguous run, found through `row_start` (first row with column >= 64 c).  yb == nullptr: off.
struct ObsEmit {
  const int* loc = nullptr;        // [n_rows] slab-local columns, ascending
  const int* row_start = nullptr;  // [n_cols / 64 + 2]
  double* yb = nullptr;
  long long ldy = 0;
  int n_rows = 0;
};

// K1  l96_forecast.cu
int l96_max_steps_per_launch();
// CTA shape code of the forecast kernel (`forced_T`): T alone = 8 points per thread; 1000 * c + T with
// c = 12 / 16: 12 / 16 points per thread, c = 17: 16 points per thread, 170-register build (T <= 192).
constexpr int l96_shape_code(int c, int T) { return 1000 * c + T; }
constexpr int l96_shape_width(int code) { return code < 1000 ? 8 * code : (code / 1000 == 12 ? 12 : 16) * (code % 1000); }
cudaError_t l96_forecast_launch(const double* in, double* out, double* traj, long long n, int k,
                                int n_steps, double dt, double F, long long in_ld, long long out_ld,
                                bool periodic, long long in_base, long long in_lo, long long in_hi,
                                long long traj_stride, int sms, int forced_T /* 0 = heuristic */,
                                cudaStream_t st, const ObsEmit* emit = nullptr);

// K1, persistent CTA-tile design  l96_forecast_v3.cu (shape code 30000 + steps per chunk)
constexpr int kL96V3Code = 30000;
bool l96_v3_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj);
long long l96_v3_scratch_ld(long long n, int n_steps);
size_t l96_v3_flag_count(long long n, int k, int n_steps);
cudaError_t l96_v3_launch(const double* in, double* out, double* scratch0, double* scratch1, long long scr_ld,
                          unsigned* flags, unsigned epoch, unsigned long long* counter, unsigned long long* ctr_base,
                          long long n, int k, int n_steps, int chunk_steps,
                          double dt, double F, long long in_ld, long long out_ld, bool periodic,
                          long long in_base, long long in_lo, long long in_hi, int sms, cudaStream_t st,
                          const ObsEmit* emit = nullptr);

// K1, chains handed out inside the SM  l96_forecast_v6.cu (shape code 60000 + variant): one persistent 384-thread CTA
// per SM; variant 0 = three 4-warp groups on 2048-point tiles, variant 1 = six 2-warp groups on 1024-point tiles; each
// group sweeps its own chain; chains of decreasing length from a shared-memory dispenser.
constexpr int kL96V6Code = 60000;
struct L96V6Plan {
  int grid = 0, variant = 0, tile_stride = 0, chain_stride = 0, hist_t = 0;
  std::vector<int> tiles;          // [grid][tile_stride][4]: {2 member + continuing, g0, p, q} in sweep order
  std::vector<int> chains;         // [grid][chain_stride]: n_chains, first tile of every chain, n_tiles
};
bool l96_v6_eligible(const double* in, const double* out, long long n, long long in_ld, long long out_ld,
                     long long in_base, int n_steps, bool traj, int variant);
bool l96_v6_plan(long long n, int k, int n_steps, int ctas, int variant, L96V6Plan* out);
cudaError_t l96_v6_launch(const double* in, double* out, const int* tiles_dev, const int* chains_dev, int grid,
                          int tile_stride, int chain_stride, int hist_t, int variant, long long n, int n_steps,
                          double dt, double F, long long in_ld, long long out_ld, bool periodic, long long in_base,
                          long long in_lo, long long in_hi, cudaStream_t st, const ObsEmit* emit = nullptr);

// Lorenz-63 members [k][3]  l63_forecast.cu
cudaError_t l63_forecast_launch(const double* x_in, double* x_out, double* traj, long long k, int n_steps,
                                double dt, double sigma, double rho, double beta, cudaStream_t st);

// Surface-QG turbulence members [k][2][N][N/2+1] complex  sqg_forecast.cu.  Every table is the caller's (host
// pointers, copied at setup): the Python mirror builds them as SQGTurb.__init__ does (dabench/data/_sqgturb.py:83-248).
struct SqgTables {
  int n = 0, ekman = 0, symmetric = 1;
  double delta_t = 0.0, inv_tdiab = 0.0, r = 0.0;
  const double *kx = nullptr, *ly = nullptr;                                  // [N/2+1], [N]
  const double *hovermu = nullptr, *tanhmu = nullptr, *sinhmu = nullptr, *ksqlsq = nullptr, *hyperdiff = nullptr;   // [N][N/2+1]
  const double* pvspec_eq = nullptr;                                          // [2][N][N/2+1] complex
};
struct SqgState;
cudaError_t sqg_setup(SqgState** out, const SqgTables& t);
void sqg_destroy(SqgState* s);
cudaError_t sqg_forecast(SqgState* s, const double* spec_in, const double* grid_in, double* spec_out, double* grid_out,
                         double* traj, int members, int n_steps, cudaStream_t st, int* n_launches);

// K2/K3  etkf_contract.cu
int contract_k8(int k);
int contract_grid(long long n_rows, int sms);
size_t contract_partial_doubles(int k, int n_cta);
cudaError_t contract_launch(const double* X, long long ld, int k, const int* obs_loc,
                            const double* obs_val, const double* obs_sw /* nullable: 1 / sigma per row */,
                            const long long* seg, int n_seg,
                            long long n_rows, double* partial, int grid,
                            bool chain /* obs tables are not the previous kernel's output: PDL allowed */,
                            cudaStream_t st);
// Multi-rank statistics exchange fused into the reduction (etkf_contract.cu): where this rank's
// reduced k*k+k values go on every rank, and the flags that announce them.
struct StatsPush {
  double* slot[kMaxRanks];               // slot `rank` of every rank's exchange buffer (this exchange's parity)
  unsigned long long* flag[kMaxRanks];   // flag `rank` on every rank (same parity)
  unsigned long long seq;                // exchange number, monotonic
  int world;                             // <= 1: no exchange, results go to `stats`
  unsigned int* done_ctr;                // local, zero: CTAs of the launch that have pushed
};
// K3 on the compact Yb the forecast kernels emit  etkf_contract_dense.cu
bool contract_dense_plan(long long n_rows, int k, int sms, int* grid_out, int* rc_out);
cudaError_t contract_dense_launch(const double* Yb, long long ldy, int k, const double* obs_val, const double* obs_sw,
                                  long long n_rows, int grid, int rc, double* partial, bool chain, cudaStream_t st);
cudaError_t stats_reduce_launch(const double* partial, int n_part, int k, double r_inv, double* stats,
                                const StatsPush* push /* nullable */, cudaStream_t st);
cudaError_t stats_gather_launch(const unsigned long long* flags, unsigned long long seq, const double* slots,
                                int k, int world, double* stats, unsigned long long timeout_ns, cudaStream_t st);
cudaError_t gather_launch(const double* X, long long ld, int k, const long long* idx,
                          const unsigned char* mask, long long n_y, double* Yb, cudaStream_t st);
cudaError_t observe_launch(const double* state, long long N, long long n_state_times, const long long* time_pos, long long n_t,
                           const long long* loc, int stationary, const long long* count, long long n_obs,
                           const double* errors, double* values, int* bad, int sms, cudaStream_t st);
cudaError_t var3d_analysis_launch(const double* xb, double* xa, double* den, long long n, const long long* loc,
                                  const double* val, long long n_rows, double w, int* bad, int sms,
                                  cudaStream_t st);
cudaError_t pack_rows_launch(const double* values, long long n_obs, long long n_times, const int* src,
                             const long long* off, const long long* src_base, int mask_nan, int* loc, double* val,
                             cudaStream_t st);
cudaError_t var3d_scatter_only_launch(const double* xb, double* num, double* den, long long n, const long long* loc,
                                      const double* val, long long n_rows, double w, int* bad, int sms, cudaStream_t st);
// var3d_cg.cu: the same analysis with a dense user B, conjugate gradients on I + B diag(w)
size_t var3d_cg_workspace_doubles(long long n);
cudaError_t var3d_cg_launch(const double* xb, double* xa, long long n, const long long* loc, const double* val,
                            long long n_rows, double w_obs, const double* B, double tol, int maxiter, double* ws,
                            int* bad, int sms, int* iters_out, long* launches, cudaStream_t st);
cudaError_t pack_obs_launch(const long long* idx, const unsigned char* mask, const double* y,
                            const double* sd /* nullable: per-row sigma */, long long n_y, long long N, int* loc,
                            double* val, double* sw /* 1 / sigma when sd != NULL */, int* bad, cudaStream_t st);

// K4  etkf_eig.cu (Jacobi eigensolver) and etkf_ns.cu (cluster Newton-Schulz, 32 < k <= 128)
enum { kEigAuto = 0, kEigJacobi = 1, kEigNewtonSchulz = 2 };
size_t eig_smem_bytes(int k);
// ns_ws: ns_ws_bytes(k) bytes of device scratch (nullable: Jacobi only).  The Newton-Schulz path is
// taken when method != kEigJacobi, k > 32, lam_out == nullptr (it has no eigenvalues to report).
cudaError_t eig_transform_launch(const double* stats, int k, double rho, int empty_R, double* T,
                                 double* lam_out, int* info_out, double* ns_ws, int method,
                                 cudaStream_t st);
size_t ns_ws_bytes(int k);
bool ns_fallback_in_kernel(int k);
cudaError_t ns_transform_launch(const double* stats, int k, double rho, double* ws, double* T,
                                int* info_out, cudaStream_t st);

// K5  etkf_transform.cu
struct TransformArgs {
  const double* T;                 // [k][k] row-major, Xa = Xb T
  const double* src[kMaxRanks];    // Xb base pointer of every rank's slab (own + peer-mapped)
  long long src_ld;                // member stride of the Xb slabs
  double* dst;                     // extended-layout output rows
  long long dst_ld, dst_base;      // member stride; column of slab position 0
  long long n_loc, n_global, n0;   // slab length, ring size, first global column of this rank
  int halo_l, halo_r, k, rank;
};
cudaError_t transform_launch(const TransformArgs& a, int sms, cudaStream_t st);
// sum over the columns of (ensemble mean - truth)^2, fixed summation order; partial: 296 doubles
cudaError_t mean_sqerr_launch(const double* xa, long long ld, int k, long long n, const double* truth,
                              double* partial, double* out, cudaStream_t st);

// etkf_batched.cu: one CTA per experiment, whole experiment in one launch
struct BatchedArgs {
  const double* x0;            // [batch][k][N]
  const double* obs_val;       // [batch][n_obs_times][n_obs] (NaN = padded slot)
  const long long* obs_idx;    // [batch or 1][n_obs_times][n_obs]
  long long idx_batch_stride;  // 0 when the index set is shared by all experiments
  long long val_batch_stride;  // 0 when the observed values are shared by all experiments
  const long long* padded;     // [n_cycles][t_pad], shared
  const double* rho;           // [batch]
  double* out_analysis;        // nullable [batch][n_cycles][k][N]
  double* out_mean;            // nullable [batch][n_cycles][N]
  double* out_final;           // nullable [batch][k][N]
  long long n_obs_times;
  int k, N, n_steps, n_cycles, t_pad, n_obs, empty_R;
  int force_jacobi;            // K4 by the Jacobi eigensolver only (dab_set_eig_method)
  double dt, F, r_inv;       // r_inv = 1 / sd^2, or 0 for sd == 0 (pinv of a zero R)
  int stationary;            // != 0: NaN observation values are data (they propagate), not padding
};
size_t batched_smem_bytes(int k, int N, int t_pad, int n_obs);
cudaError_t batched_launch(const BatchedArgs& a, int batch, cudaStream_t st);

// fp64_peak.cu: measured FP64 roofline denominators (TFLOP/s)
cudaError_t fp64_peak_measure(int sms, double* dfma_tflops, double* dmma_tflops, cudaStream_t st);

}  // namespace dab
