/* dab_b200.h -- C ABI of the B200-native ETKF cycling hot path for DataAssimBench.
 *
 * Drop-in boundary for ONE path of StevePny/DataAssimBench: `dabench.dacycler.ETKF.cycle()`
 * on a Lorenz-96 ensemble (forecast -> observation gather -> ETKF analysis).  The reference is
 * pure Python/JAX and has no FFI of its own, so each entry point below cites the reference
 * *function* it replaces (file:line under the reference repo); INTEGRATION.md shows the ctypes
 * stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain C types only; all floating data is FP64; ensembles are member-major [k][N]
 *     (the `x[ensemble, index]` array of the reference's input Dataset);
 *   - every function returns 0 on success, <0 for an argument error, >0 for a CUDA error code;
 *     the message is available from dab_last_error(); no C++ exception crosses the boundary;
 *   - "dev" pointers are device pointers owned by the caller (e.g. torch tensors); the library
 *     never frees caller memory; "host" pointers are host memory (pinned for async overlap);
 *   - work is enqueued on the stream passed in (or the handle's own stream for the cycler);
 *     nothing synchronises implicitly except dab_destroy, dab_sync and the *_host entry points;
 *   - one handle per host thread and per GPU; multi-GPU = one process (rank) per GPU.  With the peers'
 *     buffers registered (dab_cycler_ipc_import / dab_cycler_set_peer) the k*k+k statistics are summed
 *     on the device inside dab_cycler_stats (P2P stores over NVLink) and halos travel by peer-mapped
 *     loads; without them the caller all-reduces dab_cycler_stats_ptr between dab_cycler_stats and
 *     dab_cycler_update (e.g. NCCL through torch.distributed).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef DAB_B200_H
#define DAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DAB_API __attribute__((visibility("default")))
#else
#define DAB_API
#endif

typedef struct dab_handle dab_handle;

#define DAB_MAX_ENSEMBLE 128   /* k <= 128: the eigensolver lives in one CTA's shared memory */
#define DAB_MAX_RANKS 8
#define DAB_IPC_HANDLE_BYTES 64

/* ---- lifetime ------------------------------------------------------------------------- */
DAB_API int dab_version(void);
DAB_API int dab_create(dab_handle** out, int device);
DAB_API int dab_destroy(dab_handle* h);
DAB_API int dab_sync(dab_handle* h);
DAB_API const char* dab_last_error(const dab_handle* h);   /* h may be NULL: last create() error */

/* How the k x k part of the analysis (pinv + sqrtm, dabench/dacycler/_etkf.py:142-154) is solved
 * AUTO: Newton-Schulz inverse square root -- on an 8-CTA cluster for 32 < k <= 128, inside one CTA for
 * k <= 32 -- with the Jacobi eigensolver as the automatic fallback (ill-conditioned or non-finite
 * statistics, the empty-R branch, or when eigenvalues are requested); JACOBI: always the eigensolver.
 * Also settable with the environment variable DAB_EIG_METHOD=jacobi|ns before dab_create. */
#define DAB_EIG_AUTO 0
#define DAB_EIG_JACOBI 1
#define DAB_EIG_NEWTON_SCHULZ 2
DAB_API int dab_set_eig_method(dab_handle* h, int method);

/* Programmatic dependent launch between the kernels of a cycle (process-wide; default on, or the
 * environment variable DAB_PDL=0|1 read at the first launch).  With it a kernel becomes resident
 * while its predecessor drains, fetches what does not depend on it, and blocks on
 * griddepcontrol.wait before touching its output; results are bit-identical either way. */
DAB_API int dab_set_pdl(int on);

/* ---- K1: Lorenz-96 RK4 ensemble forecast -----------------------------------------------
 * Replaces ETKF._step_forecast (dabench/dacycler/_etkf.py:64-80) -> Model.forecast ->
 * Data.generate (dabench/data/_data.py:86-211) -> integrate (dabench/data/_utils.py:16-57)
 * with Lorenz96.rhs (dabench/data/_lorenz96.py:119-151), for all k members at once, with
 * fixed-step classical RK4 (BASELINE.json north star) instead of adaptive Dormand-Prince.
 *   x_in  [k][N] dev, x_out [k][N] dev (may NOT alias x_in)
 *   traj  nullable dev [n_steps][k][N]: state after step 1..n_steps
 *         (the reference's generate(n_steps+1) trajectory without its initial point) */
DAB_API int dab_l96_forecast_f64(dab_handle* h, const double* x_in, double* x_out, double* traj,
                         int64_t N, int k, int n_steps, double dt, double F, void* stream);

/* Lorenz-63 members through the same plugin point (dabench/data/_lorenz63.py:70-90 rhs; defaults
 * sigma = 10, rho = 28, beta = 8/3, :37-42), classical RK4 at dt.
 *   x_in, x_out dev [k][3] (may alias); traj nullable dev [n_steps][k][3]: state after step 1..n_steps */
DAB_API int dab_l63_forecast_f64(dab_handle* h, const double* x_in, double* x_out, double* traj, int64_t k,
                         int n_steps, double dt, double sigma, double rho, double beta, void* stream);

/* Surface quasi-geostrophic turbulence members through the same plugin point: replaces SQGTurb.integrate ->
 * lax.scan(SQGTurb._rk4) -> SQGTurb.rhs (dabench/data/_sqgturb.py:441-504, :345-364, :506-548; helpers _invert
 * :258-270, _specpad :293-309, _xyderiv_dealias :334-343, _spectrunc :311-325) for all members at once, FP64,
 * dealiased (the reference's only working branch).  The 2-D transforms on the 3N/2 grid are pruned DFTs done as
 * FP64 matrix products with all members batched (csrc/sqg_forecast.cu).
 * dab_sqg_setup copies the generator's constant tables (HOST pointers; built by the caller exactly as
 * SQGTurb.__init__ builds them, :150-248, float32 / complex64 casts included) and the DFT matrices to the device:
 *   n            grid points per direction (multiple of 4: the 3n/2 grid must be even; <= 1024); members <= 4096
 *   kx [n/2+1]   imag(ik_pad[0, :n/2+1]);  ly [n]  imag(il[:, 0]) (= il_pad on the rows _specpad fills)
 *   hovermu, tanhmu, sinhmu, ksqlsq, hyperdiff   [n][n/2+1];  pvspec_eq [2][n][n/2+1] complex128
 *   delta_t, inv_tdiab (= 1 / tdiab as the reference rounds it), r; ekman = (r >= 1e-10); symmetric
 * dab_sqg_forecast_c128 advances `members` spectral states by n_steps RK4 steps:
 *   spec_in, spec_out  dev [members][2][n][n/2+1] complex128 (re, im interleaved; may alias)
 *   traj_grid          nullable dev [n_steps][members][2][n][n] real: irfft2 of the state after step 1..n_steps
 *                      (the rows of the reference's `values`, which start one step past x0, :489-496) */
typedef struct dab_sqg_config {
  int32_t n, ekman, symmetric, reserved;
  double delta_t, inv_tdiab, r;
  const double *kx, *ly, *hovermu, *tanhmu, *sinhmu, *ksqlsq, *hyperdiff, *pvspec_eq;
} dab_sqg_config;
DAB_API int dab_sqg_setup(dab_handle* h, const dab_sqg_config* cfg);
DAB_API int dab_sqg_forecast_c128(dab_handle* h, const double* spec_in, double* spec_out, double* traj_grid,
                          int members, int n_steps, void* stream);
/* The same for GRIDDED members -- what a Model.forecast wrapper around the generator sees (state_vec in grid space,
 * SQGTurb.fft2_2dto1d in, :420-426, SQGTurb.ifft2 out, :496): rfft2 on the way in and irfft2 on the way out are two
 * more constant-matrix products, so an ensemble cycles between analysis and forecast without leaving the device.
 *   grid_in, grid_out  dev [members][2][n][n] real (may alias);  spec_out nullable: the final spectral states;
 *   traj_grid as above.  n_steps = 0 returns the projection irfft2(rfft2(grid_in)). */
DAB_API int dab_sqg_forecast_grid_f64(dab_handle* h, const double* grid_in, double* grid_out, double* spec_out,
                              double* traj_grid, int members, int n_steps, void* stream);

/* ---- K2: observation operator as a gather ----------------------------------------------
 * Replaces the dense one-hot H (DACycler._calc_default_H, dabench/dacycler/_dacycler.py:58-66),
 * its masking (dabench/dacycler/_etkf.py:202-203) and ETKF._apply_obsop `H @ Xb`
 * (dabench/dacycler/_etkf.py:82-92).  Also the device equivalent of the legacy
 * ObsOp._index_state_vec (dabench/obsop/_obsop.py:43-58).  Bit-exact copy.
 *   X [k][N] dev; idx [n_y] dev int64; mask nullable [n_y] dev uint8 (0 = row zeroed);
 *   Yb [k][n_y] dev */
DAB_API int dab_obs_gather_f64(dab_handle* h, const double* X, const int64_t* idx, const uint8_t* mask,
                       double* Yb, int64_t N, int k, int64_t n_y, void* stream);

/* ---- the step before the path: Observer.observe sampling on the device -------------------
 * Replaces the `state_vec.sel(time=...).sel(locations)` / concat / pad / `+ errors` part of
 * Observer.observe (dabench/observer/_observer.py:290-316, 346-363) for a trajectory that already
 * lives in HBM (e.g. a nature run made by dab_l96_forecast_f64): only the sampled values leave the
 * device.  The random draws (times, locations, errors) stay with the caller: they must consume
 * numpy.random.default_rng in the reference's order to stay bit-compatible.
 *   state    dev [n_state_times][N]
 *   time_pos dev int64 [n_t]            row of `state` for every observation time
 *   loc      dev int64 [n_obs] (stationary != 0) or [n_t][n_obs]
 *   count    nullable dev int64 [n_t]   valid slots per time (ragged, non-stationary); the rest is NaN
 *   errors   nullable dev [n_t][n_obs]
 *   values   dev [n_t][n_obs]           = state[time_pos[t]][loc] + errors
 * Synchronises the stream and returns -1 if a time_pos / loc entry in a valid slot is out of range. */
DAB_API int dab_observe_f64(dab_handle* h, const double* state, int64_t N, int64_t n_state_times,
                    const int64_t* time_pos, int64_t n_t, const int64_t* loc, int stationary,
                    const int64_t* count, int64_t n_obs, const double* errors, double* values,
                    void* stream);

/* ---- 3D-Var analysis with the default operators -------------------------------------------
 * Replaces Var3D._cycle_obsop (dabench/dacycler/_var3d.py:52-112) when H is the default index gather
 * (masked rows removed by the caller), R = obs_error_sd^2 I and B = I (DACycler._calc_default_B,
 * dabench/dacycler/_dacycler.py:74-76): the system I + B H^T R^-1 H the reference hands to conjugate
 * gradients is then diagonal and is solved exactly,
 *   xa[i] = (xb[i] + sum_r val[r] / sd^2) / (1 + count_i / sd^2)   over rows r with loc[r] == i.
 *   xb, xa dev [N] (distinct); loc dev int64 [n_rows]; val dev [n_rows].
 * Synchronises the stream; returns -1 if a location is out of range. */
DAB_API int dab_var3d_analysis_f64(dab_handle* h, const double* xb, double* xa, int64_t N, const int64_t* loc,
                           const double* val, int64_t n_rows, double obs_error_sd, void* stream);
/* The same analysis with a user-supplied background covariance B [N][N] (dev, row-major), the general case of
 * dabench/dacycler/_var3d.py:84-110: conjugate gradients on (I + B H^T R^-1 H) xa = xb + B H^T R^-1 y from x0 = xb,
 * stopping like jax.scipy.sparse.linalg.cg (||r||^2 <= tol^2 ||b||^2 or maxiter; the reference passes tol = 1e-5,
 * maxiter = 1000).  H^T R^-1 H is applied as a diagonal; one pass over B per iteration.  iters_out (nullable, host)
 * receives the iteration count.  Synchronises the stream. */
DAB_API int dab_var3d_analysis_b_f64(dab_handle* h, const double* xb, double* xa, int64_t N, const int64_t* loc,
                             const double* val, int64_t n_rows, double obs_error_sd, const double* B,
                             double tol, int maxiter, int* iters_out, void* stream);

/* ---- K3+K4: statistics and transform matrix (exposed for tests and for rank-parallel use) -
 * stats = [C (k*k, row-major) | g (k)] with C = Yb'^T Rinv Yb', g = Yb'^T Rinv (y - ybar)
 * (dabench/dacycler/_etkf.py:137-146,154) for R = obs_error_sd^2 I
 * (DACycler._calc_default_R, dabench/dacycler/_dacycler.py:68-72).  obs_error_sd == 0 gives Rinv = 0, as
 * the reference's pinv(R) does (stats = 0).  Synchronises the stream; returns -1 if the index of a
 * valid row is outside [0, N). */
DAB_API int dab_etkf_stats_f64(dab_handle* h, const double* X, const double* y, const int64_t* idx,
                       const uint8_t* mask, double obs_error_sd, int64_t N, int k, int64_t n_y,
                       double* stats, void* stream);
/* T [k][k] dev row-major with Xa[N,k] = Xb[N,k] T  (dabench/dacycler/_etkf.py:142-161);
 * empty_R != 0 selects the reference's zero branch (:149-152).
 * lam nullable dev [k]: eigenvalues of (k-1)/rho I + C (asking for them selects the Jacobi
 * eigensolver); sweeps nullable dev int: Jacobi sweeps or Newton-Schulz steps used. */
DAB_API int dab_etkf_transform_matrix_f64(dab_handle* h, const double* stats, int k, double rho,
                                  int empty_R, double* T, double* lam, int* sweeps, void* stream);

/* Host-only debugging aid (no device needed): the tile and chain tables of the chained forecast kernel for a shape
 * (csrc/l96_forecast_v6.cu: how the rows are cut into tiles that hand their left boundary to the next tile).
 * dims_out[4] = {grid, tile_stride, chain_stride, hist_t}; tiles_out [grid][tile_stride][4] = {2 member + continuing,
 * window start g0, first useful point p, end q}; chains_out [grid][chain_stride] = n_chains, the index of every chain's
 * first tile, n_tiles.  The tables are written only if the capacities (in ints) suffice. */
DAB_API int dab_debug_l96_chain_plan(int64_t n, int k, int n_steps, int ctas, int variant, int* dims_out, int* tiles_out,
                             int64_t tiles_cap, int* chains_out, int64_t chains_cap);

/* ---- single analysis --------------------------------------------------------------------
 * Replaces ETKF._cycle_obsop + ETKF._compute_analysis (dabench/dacycler/_etkf.py:94-213) for
 * the default operators (H = index gather, R = sd^2 I).
 *   Xb [k][N] dev, Xa [k][N] dev (may not alias), y [n_y] dev, idx [n_y] dev int64,
 *   mask nullable [n_y] dev uint8 (time mask AND location mask), rho = multiplicative inflation.
 *   n_y == 0 reproduces the reference's empty-R branch; obs_error_sd == 0 gives Rinv = 0 (pinv of a zero R:
 *   the analysis is the inflated background).  With n_y > 0 the call synchronises the stream and returns -1
 *   if the index of a valid row is outside [0, N). */
DAB_API int dab_etkf_analysis_f64(dab_handle* h, const double* Xb, double* Xa, const double* y,
                          const int64_t* idx, const uint8_t* mask, double obs_error_sd, double rho,
                          int64_t N, int k, int64_t n_y, void* stream);
/* The same for a diagonal R with unequal entries (a user R = diag(sd_r^2), dabench/dacycler/_etkf.py:142:
 * pinv of a diagonal matrix): obs_error_sd [n_y] dev, one sigma per observation; sigma == 0 gives that row
 * weight 0 (pinv).  The rows are scaled by 1 / sigma inside the contraction (K3). */
DAB_API int dab_etkf_analysis_diag_f64(dab_handle* h, const double* Xb, double* Xa, const double* y,
                               const int64_t* idx, const uint8_t* mask, const double* obs_error_sd, double rho,
                               int64_t N, int k, int64_t n_y, void* stream);

/* ---- the cycler: DACycler.cycle() for the ETKF -------------------------------------------
 * Replaces DACycler.cycle / _cycle_and_forecast (dabench/dacycler/_dacycler.py:110-141,186-290):
 * per cycle  analysis (K3,K4,K5)  then  forecast (K1), state resident in HBM throughout. */
typedef struct dab_cycle_config {
  int64_t system_dim;        /* N: global ring size */
  int32_t ensemble_dim;      /* k */
  int32_t n_steps;           /* model steps per window = steps_per_window - 1 (_dacycler.py:233) */
  double model_dt;           /* the forecast model's own delta_t */
  double forcing;            /* Lorenz-96 F */
  double rho;                /* multiplicative_inflation (_etkf.py:54) */
  double obs_error_sd;       /* R = sd^2 I */
  int64_t n_obs_times;       /* observation Dataset: sizes time x observations */
  int64_t n_obs;
  int32_t stationary;        /* attrs['stationary_observers'] (_dacycler.py:262-268) */
  int32_t n_cycles;
  int32_t t_pad;             /* columns of the padded time-index table (_utils.py:97-119) */
  int32_t rank, world;       /* grid sharding: rank owns columns [rank*N/world, (rank+1)*N/world) */
} dab_cycle_config;

/* obs_values [n_obs_times][n_obs] host (NaN = padded slot when not stationary),
 * system_index [n_obs_times][n_obs] host int64 (dabench/observer/_observer.py:327-328),
 * padded_time_idx [n_cycles][t_pad] host int64, 0 = masked, t+1 otherwise
 * (`all_filtered_padded`, dabench/dacycler/_dacycler.py:259). */
DAB_API int dab_cycler_setup(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values,
                     const int64_t* system_index, const int64_t* padded_time_idx);
/* The same with the observed values ALREADY IN HBM (obs_values_dev [n_obs_times][n_obs] dev, e.g. written by
 * dab_observe_f64): they are packed into the cycler's location-ordered table by a kernel and never visit the host;
 * the index table stays a host array (the Observer draws it on the host).  NaN slots of a moving network are masked
 * on the device. */
DAB_API int dab_cycler_setup_dev(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values_dev,
                         const int64_t* system_index, const int64_t* padded_time_idx);
/* The same with a diagonal R given slot by slot: obs_error_sd [n_obs_times][n_obs] host (cfg->obs_error_sd is
 * ignored; sigma == 0 gives the slot weight 0). */
DAB_API int dab_cycler_setup_diag(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values,
                          const int64_t* system_index, const int64_t* padded_time_idx, const double* obs_error_sd);
/* local slab [k][N/world]; is_host selects the copy direction kind */
DAB_API int dab_cycler_set_state(dab_handle* h, const double* x, int is_host);
DAB_API int dab_cycler_get_state(dab_handle* h, double* x, int is_host);
/* Run cycles [c0, c1) on one rank (world == 1).  out_analysis nullable host
 * [c1-c0][k][N]: analysis of every cycle (`isel(cycle_timestep=0)`, _dacycler.py:290), copied
 * device->host overlapped with the next cycle's compute.  out_traj nullable DEVICE
 * [c1-c0][n_steps][k][N]: forecast steps 1..n_steps of every cycle (return_forecast=True). */
DAB_API int dab_cycler_run(dab_handle* h, int c0, int c1, double* out_analysis, double* out_traj);
/* Rank-parallel phases (world >= 1): stats -> update.  The k*k+k statistics of the ranks must be
 * summed in between.  When every peer's exchange buffer (which = 2 below) is registered, that sum
 * happens ON THE DEVICE inside dab_cycler_stats -- each rank's reduction kernel stores its values
 * into all ranks' exchange buffers over NVLink and a small kernel adds them in rank order -- and
 * dab_cycler_stats_exchange() returns 1.  Otherwise (returns 0) the caller all-reduces
 * dab_cycler_stats_ptr itself (e.g. NCCL through torch.distributed) between the two calls. */
DAB_API int dab_cycler_stats(dab_handle* h, int cycle);
DAB_API int dab_cycler_stats_exchange(dab_handle* h);
DAB_API double* dab_cycler_stats_ptr(dab_handle* h);            /* dev, k*k+k doubles */
DAB_API int dab_cycler_update(dab_handle* h, int cycle, double* out_analysis_dev /* nullable [k][N/world] */);
/* Optional on-device score of every analysis (the step after the path: dabench.metrics mse/rmse of
 * the ensemble mean, dabench/metrics/_deterministic.py:43-72, without copying analyses to the
 * host): with truth_dev [n_cycles][N/world] (this rank's columns of the truth at the analysis
 * times) registered, every update also writes
 *   sse_out_dev[cycle] = sum over this rank's columns of (mean_j Xa[j][e] - truth[cycle][e])^2
 * (fixed summation order).  rmse = sqrt(sum over ranks / N).  Both NULL switches it off; the
 * pointers must stay valid until then (dab_cycler_setup also switches it off). */
DAB_API int dab_cycler_set_truth(dab_handle* h, const double* truth_dev, double* sse_out_dev);
/* Same, with this rank's slab of the analysis copied to (pinned) HOST memory on the handle's copy
 * stream, overlapped with the following cycle; dab_sync() waits for it. */
DAB_API int dab_cycler_update_host(dab_handle* h, int cycle, double* out_analysis_host /* nullable [k][N/world] */);
/* Peer plumbing for the fused halo exchange and the statistics exchange (CUDA IPC, one process per
 * GPU): which = 0/1 selects the two background (Xb) buffers, which = 2 the statistics exchange buffer. */
DAB_API int dab_cycler_ipc_export(dab_handle* h, int which, void* handle_out /* 64 bytes */);
DAB_API int dab_cycler_ipc_import(dab_handle* h, int peer_rank, int which, const void* handle_in);
/* Same-process alternative (several handles in one process, e.g. one per peer-enabled GPU, or
 * emulated ranks on one GPU in tests): register a peer's background buffer pointer directly. */
DAB_API double* dab_cycler_buffer_ptr(dab_handle* h, int which);
DAB_API int dab_cycler_set_peer(dab_handle* h, int peer_rank, int which, const double* ptr);
DAB_API void* dab_cycler_stream(dab_handle* h);                  /* cudaStream_t the cycler enqueues on */
/* device pointers of the current background / last analysis (extended layout) for tests */
DAB_API double* dab_cycler_background_ptr(dab_handle* h);
DAB_API int dab_cycler_kernel_launches(dab_handle* h);
/* What the setup-time autotuner picked for the forecast kernel: 10 * shape code + launches per window
 * (shape code: CTA width for 8 points per thread, 1000 * c + width with c = 12/16/17 for 12 / 16 /
 * 16-in-170-registers points per thread; 0 = untuned heuristic). */
DAB_API int dab_cycler_forecast_shape(dab_handle* h);           /* kernels launched since setup */
/* How many contractions since setup read the compact observation-space ensemble the forecast kernels wrote
 * (stationary networks; DAB_YB_EMIT=0 at setup switches the route off) instead of gathering H Xb from the state
 * (ETKF._apply_obsop, dabench/dacycler/_etkf.py:82-92). */
DAB_API int dab_cycler_compact_contractions(dab_handle* h);

/* Measurement aids (bench.py): per-kernel CUDA-event timing on the cycler's stream.
 * ms_out[5]/count_out[5]: contraction, reduction, eigensolver, transform, forecast. */
DAB_API int dab_cycler_set_profiling(dab_handle* h, int on);
DAB_API int dab_cycler_read_profile(dab_handle* h, double* ms_out, int* count_out);
/* FP64 roofline denominators measured on this device: independent DFMA chains and
 * DMMA.8x8x4 (TFLOP/s).  MEASURED_PEAKS.json carries no FP64 entry. */
DAB_API int dab_measure_fp64_peak(dab_handle* h, double* dfma_tflops, double* dmma_tflops);

/* One call = DACycler.cycle() with HOST buffers (world == 1): setup, H2D of the ensemble and the
 * observations, n_cycles cycles, D2H of every analysis.  x0 [k][N] host; out [n_cycles][k][N] host. */
DAB_API int dab_etkf_cycle_host_f64(dab_handle* h, const dab_cycle_config* cfg, const double* x0,
                            const double* obs_values, const int64_t* system_index,
                            const int64_t* padded_time_idx, double* out_analysis);
/* ... with a diagonal R given slot by slot (see dab_cycler_setup_diag). */
DAB_API int dab_etkf_cycle_host_diag_f64(dab_handle* h, const dab_cycle_config* cfg, const double* x0,
                                 const double* obs_values, const int64_t* system_index,
                                 const int64_t* padded_time_idx, const double* obs_error_sd, double* out_analysis);

/* Batched small problems (BASELINE.json configs[1]: 4096 x (N=40, k=20), inflation sweep): `batch`
 * independent DACycler.cycle() experiments, one CTA each, the whole experiment in ONE launch.
 * cfg gives the per-experiment shape (rho ignored; k <= 32, k*N <= 4096, t_pad*n_obs*k <= 8192).
 * All pointers are HOST memory: rho [batch]; x0 [batch][k][N]; obs_values
 * [batch][n_obs_times][n_obs] (NaN = padded slot), or [n_obs_times][n_obs] when bit 1 of
 * shared_index is set (one observation set for all experiments); system_index
 * [batch][n_obs_times][n_obs], or [n_obs_times][n_obs] when bit 0 of shared_index is set;
 * padded_time_idx [n_cycles][t_pad] (shared).
 * Outputs (each nullable): out_analysis [batch][n_cycles][k][N]; out_mean [batch][n_cycles][N]
 * (ensemble-mean analysis); out_final [batch][k][N] (background after the last forecast). */
DAB_API int dab_etkf_cycle_batched_f64(dab_handle* h, const dab_cycle_config* cfg, int batch,
                                       const double* rho, const double* x0, const double* obs_values,
                                       const int64_t* system_index, int shared_index,
                                       const int64_t* padded_time_idx, double* out_analysis,
                                       double* out_mean, double* out_final);

#ifdef __cplusplus
}
#endif
#endif /* DAB_B200_H */
