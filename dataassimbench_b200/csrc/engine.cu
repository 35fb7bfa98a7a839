// Engine + C ABI (include/dab_b200.h): owns device workspaces, streams and the cycling loop.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dab_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace dab;

namespace dab {
static int g_pdl = -1;      // -1: not decided yet (DAB_PDL, default on)
bool pdl_enabled() {
  if (g_pdl < 0) { const char* e = getenv("DAB_PDL"); g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return g_pdl == 1;
}
}  // namespace dab

int dab_set_pdl(int on) { dab::g_pdl = on ? 1 : 0; return 0; }

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

// scratch of the persistent forecast kernel (l96_forecast_v3.cu): two ping-pong chunk buffers (only touched
// when a window is cut into chunks), one flag per job (monotonic epoch, never reset), the job dispenser
// (two counters; the kernel leaves them at zero).
struct V2Work {
  DevBuf scr[2], flags, counter;
  unsigned epoch = 0;
  unsigned long long ctr_base = 0;
  cudaError_t reserve(long long n, int k, int n_steps, bool chunked) {
    const size_t bytes = chunked ? sizeof(double) * (size_t)k * (size_t)l96_v3_scratch_ld(n, n_steps) : 16;
    cudaError_t e = cudaSuccess;
    for (int w = 0; w < 2 && e == cudaSuccess; ++w) e = scr[w].reserve(bytes);
    const size_t fb = sizeof(unsigned) * l96_v3_flag_count(n, k, n_steps);
    if (e == cudaSuccess && fb > flags.bytes) {
      e = flags.reserve(fb);
      if (e == cudaSuccess) e = cudaMemset(flags.p, 0, flags.bytes);
      epoch = 0;
    }
    if (e == cudaSuccess && counter.p == nullptr) {
      e = counter.reserve(2 * sizeof(unsigned long long));
      if (e == cudaSuccess) e = cudaMemset(counter.p, 0, 2 * sizeof(unsigned long long));
      ctr_base = 0;
    }
    return e;
  }
  void release() { scr[0].release(); scr[1].release(); flags.release(); counter.release(); epoch = 0; ctr_base = 0; }
};

// tile and chain tables of the chained forecast kernel (l96_forecast_v6.cu), one per shape, built on first use
struct V6Work {
  struct Entry { long long n; int k, steps, variant, grid, tile_stride, chain_stride, hist_t; DevBuf tiles, chains; };
  std::vector<Entry*> plans;
  Entry* find(long long n, int k, int steps, int variant = 0) const {
    for (Entry* e : plans) if (e->n == n && e->k == k && e->steps == steps && e->variant == variant) return e;
    return nullptr;
  }
  void release() { for (Entry* e : plans) { e->tiles.release(); e->chains.release(); delete e; } plans.clear(); }
};

struct Cycler {
  bool ready = false;
  dab_cycle_config cfg{};
  long long n_loc = 0, n0 = 0, ld_e = 0;
  int halo_l = 0, halo_r = 0, first_steps = 0;
  int empty_R = 0;
  int cur = 0;                             // which Xb buffer holds the current background
  double* xb[2] = {nullptr, nullptr};      // plain [k][n_loc], separately cudaMalloc'ed (IPC-exportable)
  size_t xb_bytes = 0;
  DevBuf ext[2];                           // extended-layout analysis buffers
  DevBuf obs_loc, obs_val, obs_sw, seg, partial, stats, T, lam, info;
  bool has_sw = false;                     // obs_sw holds 1 / sigma per row (diagonal R with unequal entries)
  // compact observation-space ensemble written by the forecast kernel (ObsEmit, kernels.h): stationary network whose
  // every observation time has the same location-ordered rows
  DevBuf yb, row_start;
  bool emit_ok = false;                    // the network qualifies (decided at setup)
  bool yb_valid = false;                   // yb holds the observed columns of the CURRENT background
  long long emit_rows = 0, emit_loc_off = 0, ldy = 0;
  int n_compact = 0;                       // contractions since setup that read yb
  std::vector<long long> csr_off;          // [n_obs_times + 1]
  std::vector<int> n_seg;                  // per cycle
  std::vector<long long> n_rows;           // per cycle
  std::vector<long long> seg_begin;        // per cycle: offset of its rows in obs_val when it has exactly one segment, else -1
  int max_grid = 1;
  int l96_T = 0;                           // autotuned CTA width of the forecast kernel (0 = heuristic)
  int l96_chunks = 1;                      // autotuned number of launches one window is split into
  DevBuf chunk_ext[2];                     // extended-layout scratch between the chunks of a window
  const double* peer[kMaxRanks][3];        // peer-mapped buffers: Xb[0], Xb[1], statistics exchange
  bool peer_open[kMaxRanks][3];
  // statistics exchange over peer memory (world > 1): [2 parities][kMaxRanks] flags (256 bytes), then
  // [2 parities][world][k*k+k] slots; rank r writes slot r of every rank (stats_reduce_kernel)
  DevBuf xchg, xchg_ctr;
  // optional on-device score of every analysis (dab_cycler_set_truth)
  const double* truth = nullptr;           // [n_cycles][n_loc], device, caller-owned
  double* sse_out = nullptr;               // [n_cycles], device, caller-owned
  DevBuf sse_partial;
  unsigned long long xchg_seq = 0;
  cudaEvent_t ev_xa[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  bool copied_pending[2] = {false, false};
};

constexpr size_t kXchgHeader = 256;       // 2 parities x kMaxRanks 64-bit flags, padded

// the device-side statistics exchange needs every peer's exchange buffer (IPC import or set_peer)
static bool stats_exchange_ready(const Cycler& c) {
  if (c.cfg.world <= 1 || c.xchg.p == nullptr) return false;
  for (int r = 0; r < c.cfg.world; ++r)
    if (r != c.cfg.rank && c.peer[r][2] == nullptr) return false;
  return true;
}

}  // namespace

struct dab_handle {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  std::string err;
  long launches = 0;
  DevBuf s_loc, s_val, s_sw, s_seg, s_partial, s_stats, s_T;   // scratch of the standalone entry points
  DevBuf s_flag;                                         // device-side argument check of dab_observe_f64
  DevBuf s_cg;                                           // conjugate-gradient vectors of dab_var3d_analysis_b_f64
  DevBuf s_src, s_off;                                   // dab_cycler_setup_dev: row permutations and offsets
  // DAB_YB_EMIT=1 at set-up: stationary networks take the compact-Yb route (forecast kernels emit the observed
  // columns, etkf_contract_dense.cu contracts them).  Off by default: at C3 the contraction gains 3.5 us and the
  // forecast loses 4.5 (profiles/yb_emit_r02.md); larger shapes do not fit a CTA's rows in shared memory.
  bool yb_emit = false;
  DevBuf ns_ws;                                          // Newton-Schulz scratch of K4 (L2-resident)
  V2Work v2;                                             // persistent forecast kernel: scratch, flags, dispenser
  V6Work v6;                                             // chained forecast kernel: tile and chain tables
  SqgState* sqg = nullptr;                               // surface-QG generator: tables, DFT matrices, workspaces
  int eig_method = kEigAuto;
  unsigned long long xchg_timeout_ns = 10000000000ull;   // statistics exchange: give up on a lost peer (DAB_XCHG_TIMEOUT_S, 0 = never)
  DevBuf b_x0, b_val, b_idx, b_pad, b_rho, b_out, b_mean, b_fin;   // batched small-problem path
  Cycler cyc;
  // optional per-kernel timing (CUDA events on the launching stream)
  // forecast autotune cache: (n_loc, k, steps) -> CTA width
  std::vector<long long> tune_key;
  std::vector<int> tune_T, tune_chunks;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<int> prof_tags;
  size_t prof_used = 0;
};

enum { PROF_START = -1, PROF_CONTRACT = 0, PROF_REDUCE = 1, PROF_EIG = 2, PROF_TRANSFORM = 3, PROF_FORECAST = 4 };

static void prof_mark(dab_handle* h, int tag) {
  if (!h->profiling) return;
  if (h->prof_used == h->prof_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->prof_pool.push_back(e);
  }
  cudaEventRecord(h->prof_pool[h->prof_used], h->stream);
  if (h->prof_tags.size() <= h->prof_used) h->prof_tags.resize(h->prof_used + 1);
  h->prof_tags[h->prof_used] = tag;
  h->prof_used++;
}

// R^-1 for R = sd^2 I the way the reference gets it: pinv(R, rtol=1e-15) (dabench/dacycler/_etkf.py:142),
// so sd == 0 gives Rinv = 0 (the observations carry no weight; the analysis is the inflated background)
// and not infinity.  Observer's default error_sd is 0.0, so this case is reachable by default.
static double r_inv_of(double sd) { return sd == 0.0 ? 0.0 : 1.0 / (sd * sd); }

static std::string g_create_err;

static int fail(dab_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (h) h->err = buf; else g_create_err = buf;
  return code;
}

#define DAB_CUDA(h, expr)                                                                    \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail((h), (int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define DAB_ARG(h, cond)                                                                     \
  do { if (!(cond)) return fail((h), -1, "invalid argument: %s", #cond); } while (0)

static void cycler_close_peers(dab_handle* h) {
  Cycler& c = h->cyc;
  for (int r = 0; r < kMaxRanks; ++r)
    for (int w = 0; w < 3; ++w) {
      if (c.peer_open[r][w]) { cudaIpcCloseMemHandle(const_cast<double*>(c.peer[r][w])); c.peer_open[r][w] = false; }
      c.peer[r][w] = nullptr;
    }
}

// Workspaces are grow-only and survive dab_cycler_setup calls (cudaMalloc/cudaFree cost
// milliseconds and synchronise the device); everything is released in dab_destroy.
static void cycler_free(dab_handle* h) {
  Cycler& c = h->cyc;
  cycler_close_peers(h);
  for (int w = 0; w < 2; ++w) {
    if (c.xb[w]) cudaFree(c.xb[w]);
    c.xb[w] = nullptr;
    c.ext[w].release();
    c.chunk_ext[w].release();
    if (c.ev_xa[w]) cudaEventDestroy(c.ev_xa[w]);
    if (c.ev_copied[w]) cudaEventDestroy(c.ev_copied[w]);
    c.ev_xa[w] = c.ev_copied[w] = nullptr;
    c.copied_pending[w] = false;
  }
  c.xb_bytes = 0;
  c.obs_loc.release(); c.obs_val.release(); c.obs_sw.release(); c.seg.release(); c.partial.release();
  c.yb.release(); c.row_start.release();
  c.stats.release(); c.T.release(); c.lam.release(); c.info.release();
  c.xchg.release(); c.xchg_ctr.release(); c.sse_partial.release();
  c.ready = false;
}

// one launch of the persistent forecast kernel (workspace reserved by the caller)
static cudaError_t forecast_v3(dab_handle* h, const double* in, double* out, long long n, int k, int steps,
                               int chunk_steps, double dt, double F, long long in_ld, long long out_ld,
                               bool periodic, long long in_base, long long in_lo, long long in_hi,
                               const ObsEmit* emit = nullptr) {
  V2Work& w = h->v2;
  if (++w.epoch == 0) {                       // 2^32 launches: start the flags over
    cudaError_t e = cudaMemsetAsync(w.flags.p, 0, w.flags.bytes, h->stream);
    if (e != cudaSuccess) return e;
    w.epoch = 1;
  }
  return l96_v3_launch(in, out, w.scr[0].as<double>(), w.scr[1].as<double>(),
                       periodic ? n : l96_v3_scratch_ld(n, steps), w.flags.as<unsigned>(), w.epoch,
                       w.counter.as<unsigned long long>(), &w.ctr_base, n, k, steps, chunk_steps, dt, F, in_ld,
                       out_ld, periodic, in_base, in_lo, in_hi, h->sms, h->stream, emit);
}

// one launch of the chained forecast kernel; the tables of a shape are planned and uploaded on first use
static cudaError_t forecast_v6(dab_handle* h, int variant, const double* in, double* out, long long n, int k, int steps,
                               double dt, double F, long long in_ld, long long out_ld, bool periodic,
                               long long in_base, long long in_lo, long long in_hi, cudaStream_t st,
                               const ObsEmit* emit = nullptr) {
  V6Work::Entry* e = h->v6.find(n, k, steps, variant);
  if (e == nullptr) {
    L96V6Plan plan;
    if (!l96_v6_plan(n, k, steps, h->sms, variant, &plan)) return cudaErrorInvalidValue;
    e = new V6Work::Entry();
    e->n = n; e->k = k; e->steps = steps; e->variant = variant;
    e->grid = plan.grid; e->tile_stride = plan.tile_stride; e->chain_stride = plan.chain_stride; e->hist_t = plan.hist_t;
    cudaError_t ce = e->tiles.reserve(plan.tiles.size() * sizeof(int));
    if (ce == cudaSuccess) ce = e->chains.reserve(plan.chains.size() * sizeof(int));
    // ordered on the launching stream, then waited for: a plain cudaMemcpy from pageable memory may return before
    // its DMA has landed, and `st` (non-blocking) is not ordered after the default stream
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->tiles.p, plan.tiles.data(), plan.tiles.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->chains.p, plan.chains.data(), plan.chains.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { e->tiles.release(); e->chains.release(); delete e; return ce; }
    h->v6.plans.push_back(e);
  }
  return l96_v6_launch(in, out, e->tiles.as<int>(), e->chains.as<int>(), e->grid, e->tile_stride, e->chain_stride, e->hist_t,
                       variant, n, steps, dt, F, in_ld, out_ld, periodic, in_base, in_lo, in_hi, st, emit);
}

// K1 over one window of `steps` <= first_steps steps: extended-layout analysis `in_ext` -> `dst`
// ([k][n_loc]).  chunks > 1 splits the window into shorter launches through extended scratch
// buffers: the tile halo is 12 points per step OF THE LAUNCH, so shorter launches recompute less
// (useful fraction 88 % -> 94 % at 20 -> 10 steps, 2048-point tiles) and quantise into waves more
// finely, at the price of one more pass over the (L2- or HBM-resident) state per chunk.  Chunk j
// with `rem` steps still to go writes the slab plus the 8 rem / 4 rem halo the rest needs.
// `emit` (nullable): also write the observed columns of the result into the compact Yb; *emitted says whether the
// launch form used could do it (a window cut into several launches does not).
static cudaError_t forecast_ext(dab_handle* h, Cycler& c, const double* in_ext, double* dst, double* traj,
                                long long traj_stride, int steps, int T, int chunks, int* n_launches,
                                const ObsEmit* emit = nullptr, bool* emitted = nullptr) {
  const int k = c.cfg.ensemble_dim;
  if (emitted) *emitted = false;
  if (T >= kL96V6Code && l96_v6_eligible(in_ext, dst, c.n_loc, c.ld_e, c.n_loc, c.halo_l, steps, traj != nullptr, T - kL96V6Code)) {
    cudaError_t e = forecast_v6(h, T - kL96V6Code, in_ext, dst, c.n_loc, k, steps, c.cfg.model_dt, c.cfg.forcing, c.ld_e, c.n_loc, false,
                                c.halo_l, -8ll * steps, c.n_loc + 4ll * steps, h->stream, emit);
    if (e == cudaSuccess && n_launches) ++*n_launches;
    if (e == cudaSuccess && emitted) *emitted = emit != nullptr;
    return e;
  }
  if (T >= kL96V6Code) T = 0;
  if (T >= kL96V3Code && l96_v3_eligible(in_ext, dst, c.n_loc, c.ld_e, c.n_loc, c.halo_l, steps, traj != nullptr)) {
    const int chunk_steps = T - kL96V3Code;
    cudaError_t e = h->v2.reserve(c.n_loc, k, c.first_steps, chunk_steps < steps);
    if (e != cudaSuccess) return e;
    e = forecast_v3(h, in_ext, dst, c.n_loc, k, steps, chunk_steps, c.cfg.model_dt, c.cfg.forcing, c.ld_e, c.n_loc,
                    false, c.halo_l, -8ll * steps, c.n_loc + 4ll * steps, emit);
    if (e == cudaSuccess && n_launches) ++*n_launches;
    if (e == cudaSuccess && emitted) *emitted = emit != nullptr;
    return e;
  }
  if (T >= kL96V3Code) T = 0;                 // not eligible (odd sizes, trajectory wanted): first design, heuristic shape
  if (traj != nullptr || chunks <= 1 || steps < 2 * chunks) chunks = 1;
  const double* src = in_ext;
  int done = 0;
  for (int j = 0; j < chunks; ++j) {
    const int sj = (steps - done + (chunks - j) - 1) / (chunks - j);
    const int rem = steps - done - sj;
    const long long n2 = c.n_loc + 12ll * rem;
    double* out = dst;
    long long out_ld = c.n_loc;
    if (rem > 0) {
      cudaError_t e = c.chunk_ext[j & 1].reserve(sizeof(double) * (size_t)k * c.ld_e);
      if (e != cudaSuccess) return e;
      out = c.chunk_ext[j & 1].as<double>() + (c.halo_l - 8 * rem);
      out_ld = c.ld_e;
    }
    const bool last = rem == 0;
    cudaError_t e = l96_forecast_launch(src, out, traj, n2, k, sj, c.cfg.model_dt, c.cfg.forcing, c.ld_e, out_ld,
                                        false, c.halo_l - 8ll * rem, -8ll * sj, n2 + 4ll * sj, traj_stride,
                                        h->sms, T, h->stream, last ? emit : nullptr);
    if (e != cudaSuccess) return e;
    if (last && emitted) *emitted = emit != nullptr;
    if (n_launches) ++*n_launches;
    src = c.chunk_ext[j & 1].as<double>();
    done += sj;
  }
  return cudaSuccess;
}

extern "C" {

int dab_version(void) { return 100; }

int dab_create(dab_handle** out, int device) {
  if (!out) return fail(nullptr, -1, "dab_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice,
                "dab_create: no CUDA device (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, -1, "dab_create: device %d out of range [0,%d)", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, (int)e, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return fail(nullptr, (int)e, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major < 10)
    return fail(nullptr, -2, "dab_create: device %s is sm_%d%d; this library is built for sm_100a only",
                p.name, p.major, p.minor);
  dab_handle* h = new dab_handle();
  h->device = device;
  h->sms = p.multiProcessorCount;
  memset(h->cyc.peer, 0, sizeof(h->cyc.peer));
  memset(h->cyc.peer_open, 0, sizeof(h->cyc.peer_open));
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = h->ns_ws.reserve(ns_ws_bytes(kMaxEns));
  if (e == cudaSuccess) e = cudaMemset(h->ns_ws.p, 0, 16);
  if (e != cudaSuccess) { delete h; return fail(nullptr, (int)e, "dab_create: %s", cudaGetErrorString(e)); }
  if (const char* ev = getenv("DAB_XCHG_TIMEOUT_S")) { const double v = atof(ev); h->xchg_timeout_ns = v > 0.0 ? (unsigned long long)(v * 1e9) : 0ull; }
  if (const char* ev = getenv("DAB_EIG_METHOD")) {
    if (!strcmp(ev, "jacobi")) h->eig_method = kEigJacobi;
    else if (!strcmp(ev, "ns")) h->eig_method = kEigNewtonSchulz;
  }
  *out = h;
  return 0;
}

int dab_destroy(dab_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cycler_free(h);
  h->s_loc.release(); h->s_val.release(); h->s_sw.release(); h->s_seg.release();
  h->s_partial.release(); h->s_stats.release(); h->s_T.release(); h->ns_ws.release(); h->s_flag.release(); h->s_cg.release(); h->s_src.release(); h->s_off.release();
  h->v2.release();
  h->v6.release();
  sqg_destroy(h->sqg); h->sqg = nullptr;
  h->b_x0.release(); h->b_val.release(); h->b_idx.release(); h->b_pad.release(); h->b_rho.release();
  h->b_out.release(); h->b_mean.release(); h->b_fin.release();
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return 0;
}

int dab_sync(dab_handle* h) {
  DAB_ARG(h, h != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, cudaStreamSynchronize(h->stream));
  DAB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  return 0;
}

int dab_set_eig_method(dab_handle* h, int method) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, method == DAB_EIG_AUTO || method == DAB_EIG_JACOBI || method == DAB_EIG_NEWTON_SCHULZ);
  h->eig_method = method;
  return 0;
}

const char* dab_last_error(const dab_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

// ------------------------------------------------------------------------------------ K1
static int forecast_plain(dab_handle* h, const double* x_in, double* x_out, double* traj,
                          long long N, int k, int n_steps, double dt, double F, cudaStream_t st,
                          double* tmp /* [k][N] scratch, needed when chunking */) {
  const int maxs = l96_max_steps_per_launch();
  const long long traj_stride = (long long)k * N;
  if (n_steps == 0) {
    DAB_CUDA(h, cudaMemcpyAsync(x_out, x_in, sizeof(double) * k * N, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const int n_chunks = (n_steps + maxs - 1) / maxs;
  const double* src = x_in;
  int done = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const int s = (n_steps - done) < maxs ? (n_steps - done) : maxs;
    // ping-pong so that the last chunk lands in x_out
    double* dst = ((n_chunks - 1 - c) % 2 == 0) ? x_out : tmp;
    DAB_CUDA(h, l96_forecast_launch(src, dst, traj ? traj + (long long)done * traj_stride : nullptr, N, k, s,
                                    dt, F, N, N, true, 0, 0, 0, traj_stride, h->sms, 0, st));
    h->launches++;
    src = dst;
    done += s;
  }
  return 0;
}

// Host-only: the tile / chain tables the chained forecast kernel (l96_forecast_v6.cu) would be launched with, so that
// the tiling logic can be checked without a GPU (tests/test_cabi_cpu.py).  dims_out = {grid, tile_stride, chain_stride,
// hist_t}; tiles_out / chains_out receive the tables if they are large enough (sizes in ints), else only dims_out.
int dab_debug_l96_chain_plan(int64_t n, int k, int n_steps, int ctas, int variant, int* dims_out, int* tiles_out,
                             int64_t tiles_cap, int* chains_out, int64_t chains_cap) {
  if (!dims_out || n < 4 || k < 1) return -1;
  L96V6Plan plan;
  if (!l96_v6_plan(n, k, n_steps, ctas, variant, &plan)) return -1;
  dims_out[0] = plan.grid; dims_out[1] = plan.tile_stride; dims_out[2] = plan.chain_stride; dims_out[3] = plan.hist_t;
  if (tiles_out && (int64_t)plan.tiles.size() <= tiles_cap) memcpy(tiles_out, plan.tiles.data(), plan.tiles.size() * sizeof(int));
  if (chains_out && (int64_t)plan.chains.size() <= chains_cap) memcpy(chains_out, plan.chains.data(), plan.chains.size() * sizeof(int));
  return 0;
}

int dab_l96_forecast_f64(dab_handle* h, const double* x_in, double* x_out, double* traj,
                         int64_t N, int k, int n_steps, double dt, double F, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, x_in != nullptr && x_out != nullptr && x_in != x_out);
  DAB_ARG(h, N >= 4 && k >= 1 && n_steps >= 0);
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* tmp = nullptr;
  if (n_steps > l96_max_steps_per_launch()) {
    DAB_CUDA(h, h->s_partial.reserve(sizeof(double) * k * N));
    tmp = h->s_partial.as<double>();
  }
  return forecast_plain(h, x_in, x_out, traj, N, k, n_steps, dt, F, st, tmp);
}

// Lorenz-63: members are [k][3]; in place (x_in == x_out) is allowed, every thread owns one member
int dab_l63_forecast_f64(dab_handle* h, const double* x_in, double* x_out, double* traj, int64_t k,
                         int n_steps, double dt, double sigma, double rho, double beta, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, k >= 0 && n_steps >= 0);
  DAB_ARG(h, k == 0 || (x_in != nullptr && x_out != nullptr));
  if (k == 0) return 0;
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, l63_forecast_launch(x_in, x_out, traj, k, n_steps, dt, sigma, rho, beta,
                                  static_cast<cudaStream_t>(stream)));
  h->launches++;
  return 0;
}

int dab_sqg_setup(dab_handle* h, const dab_sqg_config* cfg) {
  DAB_ARG(h, h != nullptr && cfg != nullptr);
  DAB_ARG(h, cfg->n >= 4 && cfg->n % 4 == 0 && cfg->n <= 1024);
  DAB_ARG(h, cfg->kx && cfg->ly && cfg->hovermu && cfg->tanhmu && cfg->sinhmu && cfg->ksqlsq && cfg->hyperdiff && cfg->pvspec_eq);
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, cudaDeviceSynchronize());
  sqg_destroy(h->sqg); h->sqg = nullptr;
  SqgTables t;
  t.n = cfg->n; t.ekman = cfg->ekman; t.symmetric = cfg->symmetric;
  t.delta_t = cfg->delta_t; t.inv_tdiab = cfg->inv_tdiab; t.r = cfg->r;
  t.kx = cfg->kx; t.ly = cfg->ly; t.hovermu = cfg->hovermu; t.tanhmu = cfg->tanhmu; t.sinhmu = cfg->sinhmu;
  t.ksqlsq = cfg->ksqlsq; t.hyperdiff = cfg->hyperdiff; t.pvspec_eq = cfg->pvspec_eq;
  DAB_CUDA(h, sqg_setup(&h->sqg, t));
  return 0;
}

int dab_sqg_forecast_c128(dab_handle* h, const double* spec_in, double* spec_out, double* traj_grid, int members,
                          int n_steps, void* stream) {
  DAB_ARG(h, h != nullptr);
  if (h->sqg == nullptr) return fail(h, -1, "dab_sqg_forecast_c128: call dab_sqg_setup first");
  DAB_ARG(h, members >= 0 && members <= 4096 && n_steps >= 0);      // (member, field) is a grid dimension
  DAB_ARG(h, members == 0 || (spec_in != nullptr && spec_out != nullptr));
  if (members == 0) return 0;
  DAB_CUDA(h, cudaSetDevice(h->device));
  int nl = 0;
  DAB_CUDA(h, sqg_forecast(h->sqg, spec_in, nullptr, spec_out, nullptr, traj_grid, members, n_steps, static_cast<cudaStream_t>(stream), &nl));
  h->launches += nl;
  return 0;
}

int dab_sqg_forecast_grid_f64(dab_handle* h, const double* grid_in, double* grid_out, double* spec_out, double* traj_grid,
                              int members, int n_steps, void* stream) {
  DAB_ARG(h, h != nullptr);
  if (h->sqg == nullptr) return fail(h, -1, "dab_sqg_forecast_grid_f64: call dab_sqg_setup first");
  DAB_ARG(h, members >= 0 && members <= 4096 && n_steps >= 0);
  DAB_ARG(h, members == 0 || (grid_in != nullptr && grid_out != nullptr));
  if (members == 0) return 0;
  DAB_CUDA(h, cudaSetDevice(h->device));
  int nl = 0;
  DAB_CUDA(h, sqg_forecast(h->sqg, nullptr, grid_in, spec_out, grid_out, traj_grid, members, n_steps, static_cast<cudaStream_t>(stream), &nl));
  h->launches += nl;
  return 0;
}

// ------------------------------------------------------------------------------------ K2
int dab_obs_gather_f64(dab_handle* h, const double* X, const int64_t* idx, const uint8_t* mask,
                       double* Yb, int64_t N, int k, int64_t n_y, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, N >= 1 && k >= 1 && n_y >= 0);
  DAB_ARG(h, n_y == 0 || (X != nullptr && idx != nullptr && Yb != nullptr));
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, gather_launch(X, N, k, reinterpret_cast<const long long*>(idx), mask, n_y, Yb,
                            static_cast<cudaStream_t>(stream)));
  if (n_y > 0) h->launches++;
  return 0;
}

// ------------------------------------------------------------- Observer.observe sampling
int dab_observe_f64(dab_handle* h, const double* state, int64_t N, int64_t n_state_times,
                    const int64_t* time_pos, int64_t n_t, const int64_t* loc, int stationary,
                    const int64_t* count, int64_t n_obs, const double* errors, double* values,
                    void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, N >= 1 && n_state_times >= 1 && n_t >= 0 && n_obs >= 0);
  DAB_ARG(h, n_t * n_obs == 0 || (state != nullptr && time_pos != nullptr && loc != nullptr && values != nullptr));
  if (n_t * n_obs == 0) return 0;
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DAB_CUDA(h, h->s_flag.reserve(sizeof(int)));
  DAB_CUDA(h, cudaMemsetAsync(h->s_flag.p, 0, sizeof(int), st));
  DAB_CUDA(h, observe_launch(state, N, n_state_times, reinterpret_cast<const long long*>(time_pos), n_t,
                             reinterpret_cast<const long long*>(loc), stationary,
                             reinterpret_cast<const long long*>(count), n_obs, errors, values,
                             h->s_flag.as<int>(), h->sms, st));
  h->launches++;
  int bad = 0;
  DAB_CUDA(h, cudaMemcpyAsync(&bad, h->s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  DAB_CUDA(h, cudaStreamSynchronize(st));
  if (bad) return fail(h, -1, "dab_observe_f64: a time_pos or loc entry of a valid slot is out of range");
  return 0;
}

// ------------------------------------------------------- 3D-Var analysis, default operators
int dab_var3d_analysis_f64(dab_handle* h, const double* xb, double* xa, int64_t N, const int64_t* loc,
                           const double* val, int64_t n_rows, double obs_error_sd, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, N >= 1 && n_rows >= 0 && obs_error_sd > 0.0);
  DAB_ARG(h, xb != nullptr && xa != nullptr && xb != xa);
  DAB_ARG(h, n_rows == 0 || (loc != nullptr && val != nullptr));
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DAB_CUDA(h, h->s_flag.reserve(sizeof(int)));
  DAB_CUDA(h, h->s_val.reserve(sizeof(double) * (size_t)N));          // the denominators
  DAB_CUDA(h, cudaMemsetAsync(h->s_flag.p, 0, sizeof(int), st));
  DAB_CUDA(h, var3d_analysis_launch(xb, xa, h->s_val.as<double>(), N, reinterpret_cast<const long long*>(loc),
                                    val, n_rows, 1.0 / (obs_error_sd * obs_error_sd), h->s_flag.as<int>(),
                                    h->sms, st));
  h->launches += n_rows > 0 ? 3 : 2;
  int bad = 0;
  DAB_CUDA(h, cudaMemcpyAsync(&bad, h->s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  DAB_CUDA(h, cudaStreamSynchronize(st));
  if (bad) return fail(h, -1, "dab_var3d_analysis_f64: an observation location is out of range");
  return 0;
}

int dab_var3d_analysis_b_f64(dab_handle* h, const double* xb, double* xa, int64_t N, const int64_t* loc,
                             const double* val, int64_t n_rows, double obs_error_sd, const double* B,
                             double tol, int maxiter, int* iters_out, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, N >= 1 && n_rows >= 0 && obs_error_sd > 0.0 && tol >= 0.0 && maxiter >= 0);
  DAB_ARG(h, xb != nullptr && xa != nullptr && xb != xa && B != nullptr);
  DAB_ARG(h, n_rows == 0 || (loc != nullptr && val != nullptr));
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DAB_CUDA(h, h->s_flag.reserve(sizeof(int)));
  DAB_CUDA(h, h->s_cg.reserve(sizeof(double) * var3d_cg_workspace_doubles(N)));
  DAB_CUDA(h, cudaMemsetAsync(h->s_flag.p, 0, sizeof(int), st));
  int iters = 0;
  DAB_CUDA(h, var3d_cg_launch(xb, xa, N, reinterpret_cast<const long long*>(loc), val, n_rows,
                              1.0 / (obs_error_sd * obs_error_sd), B, tol, maxiter, h->s_cg.as<double>(),
                              h->s_flag.as<int>(), h->sms, &iters, &h->launches, st));
  if (iters_out) *iters_out = iters;
  int bad = 0;
  DAB_CUDA(h, cudaMemcpyAsync(&bad, h->s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  DAB_CUDA(h, cudaStreamSynchronize(st));
  if (bad) return fail(h, -1, "dab_var3d_analysis_b_f64: an observation location is out of range");
  return 0;
}

// ------------------------------------------------------------------------------- K3 / K4
// `sd_vec` (nullable, device, [n_y]): per-observation sigma of a diagonal R; `sd` is then ignored
static int stats_from_arrays(dab_handle* h, const double* X, const double* y, const int64_t* idx,
                             const uint8_t* mask, double sd, const double* sd_vec, long long N, int k, long long n_y,
                             double* stats, cudaStream_t st) {
  DAB_CUDA(h, h->s_loc.reserve(sizeof(int) * (size_t)(n_y > 0 ? n_y : 1)));
  DAB_CUDA(h, h->s_val.reserve(sizeof(double) * (size_t)(n_y > 0 ? n_y : 1)));
  if (sd_vec) DAB_CUDA(h, h->s_sw.reserve(sizeof(double) * (size_t)(n_y > 0 ? n_y : 1)));
  DAB_CUDA(h, h->s_seg.reserve(sizeof(long long) * 3));
  const int grid = contract_grid(n_y, h->sms);
  DAB_CUDA(h, h->s_partial.reserve(sizeof(double) * contract_partial_doubles(k, grid)));
  DAB_CUDA(h, h->s_flag.reserve(sizeof(int)));
  DAB_CUDA(h, cudaMemsetAsync(h->s_flag.p, 0, sizeof(int), st));
  DAB_CUDA(h, pack_obs_launch(reinterpret_cast<const long long*>(idx), mask, y, sd_vec, n_y, N,
                              h->s_loc.as<int>(), h->s_val.as<double>(), h->s_sw.as<double>(), h->s_flag.as<int>(), st));
  const long long seg[3] = {0, n_y, 0};
  DAB_CUDA(h, cudaMemcpyAsync(h->s_seg.p, seg, sizeof(seg), cudaMemcpyHostToDevice, st));
  DAB_CUDA(h, contract_launch(X, N, k, h->s_loc.as<int>(), h->s_val.as<double>(), sd_vec ? h->s_sw.as<double>() : nullptr,
                              h->s_seg.as<long long>(), 1, n_y, h->s_partial.as<double>(), grid, false, st));
  DAB_CUDA(h, stats_reduce_launch(h->s_partial.as<double>(), grid, k, sd_vec ? 1.0 : r_inv_of(sd), stats, nullptr, st));
  h->launches += (n_y > 0 ? 1 : 0) + 2;
  return 0;
}

// the device-side index check of pack_obs_kernel: read back (synchronises the stream)
static int check_index_flag(dab_handle* h, const char* who, cudaStream_t st) {
  int bad = 0;
  DAB_CUDA(h, cudaMemcpyAsync(&bad, h->s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  DAB_CUDA(h, cudaStreamSynchronize(st));
  if (bad) return fail(h, -1, "%s: an observation index of a valid row is outside [0, N)", who);
  return 0;
}

int dab_etkf_stats_f64(dab_handle* h, const double* X, const double* y, const int64_t* idx,
                       const uint8_t* mask, double obs_error_sd, int64_t N, int k, int64_t n_y,
                       double* stats, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, X != nullptr && stats != nullptr && N >= 1 && n_y >= 0);
  DAB_ARG(h, k >= 2 && k <= DAB_MAX_ENSEMBLE);
  DAB_ARG(h, obs_error_sd == obs_error_sd);            // not NaN; 0 means Rinv = 0 (see r_inv_of)
  DAB_ARG(h, N < (1ll << 31));
  DAB_ARG(h, n_y == 0 || (y != nullptr && idx != nullptr));
  DAB_CUDA(h, cudaSetDevice(h->device));
  int rc = stats_from_arrays(h, X, y, idx, mask, obs_error_sd, nullptr, N, k, n_y, stats,
                             static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return n_y > 0 ? check_index_flag(h, "dab_etkf_stats_f64", static_cast<cudaStream_t>(stream)) : 0;
}

int dab_etkf_transform_matrix_f64(dab_handle* h, const double* stats, int k, double rho,
                                  int empty_R, double* T, double* lam, int* sweeps, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, (stats != nullptr || empty_R) && T != nullptr);
  DAB_ARG(h, k >= 2 && k <= DAB_MAX_ENSEMBLE && rho > 0.0);
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, eig_transform_launch(stats, k, rho, empty_R, T, lam, sweeps, h->ns_ws.as<double>(), h->eig_method,
                                   static_cast<cudaStream_t>(stream)));
  h->launches++;
  return 0;
}

static int analysis_impl(dab_handle* h, const double* Xb, double* Xa, const double* y,
                         const int64_t* idx, const uint8_t* mask, double obs_error_sd, const double* sd_vec, double rho,
                         int64_t N, int k, int64_t n_y, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, Xb != nullptr && Xa != nullptr && Xb != Xa);
  DAB_ARG(h, N >= 1 && n_y >= 0 && k >= 2 && k <= DAB_MAX_ENSEMBLE);
  DAB_ARG(h, rho > 0.0 && obs_error_sd == obs_error_sd);
  DAB_ARG(h, n_y == 0 || (y != nullptr && idx != nullptr));
  DAB_ARG(h, N < (1ll << 31));
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DAB_CUDA(h, h->s_stats.reserve(sizeof(double) * ((size_t)k * k + k)));
  DAB_CUDA(h, h->s_T.reserve(sizeof(double) * (size_t)k * k));
  const int empty_R = n_y == 0;
  if (!empty_R) {
    int rc = stats_from_arrays(h, Xb, y, idx, mask, obs_error_sd, sd_vec, N, k, n_y, h->s_stats.as<double>(), st);
    if (rc) return rc;
  }
  DAB_CUDA(h, eig_transform_launch(h->s_stats.as<double>(), k, rho, empty_R, h->s_T.as<double>(),
                                   nullptr, nullptr, h->ns_ws.as<double>(), h->eig_method, st));
  TransformArgs a{};
  a.T = h->s_T.as<double>();
  a.src[0] = Xb; a.src_ld = N;
  a.dst = Xa; a.dst_ld = N; a.dst_base = 0;
  a.n_loc = N; a.n_global = N; a.n0 = 0;
  a.halo_l = 0; a.halo_r = 0; a.k = k; a.rank = 0;
  DAB_CUDA(h, transform_launch(a, h->sms, st));
  h->launches += 2;
  return empty_R ? 0 : check_index_flag(h, "dab_etkf_analysis_f64", st);
}

int dab_etkf_analysis_f64(dab_handle* h, const double* Xb, double* Xa, const double* y,
                          const int64_t* idx, const uint8_t* mask, double obs_error_sd, double rho,
                          int64_t N, int k, int64_t n_y, void* stream) {
  return analysis_impl(h, Xb, Xa, y, idx, mask, obs_error_sd, nullptr, rho, N, k, n_y, stream);
}

int dab_etkf_analysis_diag_f64(dab_handle* h, const double* Xb, double* Xa, const double* y,
                               const int64_t* idx, const uint8_t* mask, const double* obs_error_sd, double rho,
                               int64_t N, int k, int64_t n_y, void* stream) {
  DAB_ARG(h, h != nullptr);
  DAB_ARG(h, n_y == 0 || obs_error_sd != nullptr);
  return analysis_impl(h, Xb, Xa, y, idx, mask, 1.0, obs_error_sd, rho, N, k, n_y, stream);
}

// ------------------------------------------------------------------------------ the cycler
// `x0_host` (nullable): initial ensemble in host memory; its upload is enqueued as soon as the
// buffers exist, so that it overlaps the host-side preparation of the observation tables.
// `obs_sd` (nullable, host, [n_obs_times][n_obs]): per-slot sigma of a diagonal R; cfg->obs_error_sd is then ignored.
// `obs_values_dev` (nullable, device, [n_obs_times][n_obs]): the observed values already in HBM (an Observer that
// sampled a device trajectory): `obs_values` is then not read, the location-ordered value table is filled by a
// kernel from the permutation the index rows give, and NaN slots of a moving network are masked on the device.
static int cycler_setup_impl(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values,
                             const int64_t* system_index, const int64_t* padded_time_idx,
                             const double* x0_host, const double* obs_sd = nullptr,
                             const double* obs_values_dev = nullptr) {
  DAB_ARG(h, h != nullptr && cfg != nullptr);
  DAB_ARG(h, cfg->system_dim >= 4 && cfg->system_dim < (1ll << 31));
  DAB_ARG(h, cfg->ensemble_dim >= 2 && cfg->ensemble_dim <= DAB_MAX_ENSEMBLE);
  DAB_ARG(h, cfg->n_steps >= 0 && cfg->n_cycles >= 0 && cfg->t_pad >= 0);
  DAB_ARG(h, cfg->rho > 0.0 && cfg->model_dt > 0.0);
  DAB_ARG(h, cfg->obs_error_sd == cfg->obs_error_sd);   // not NaN; 0 means Rinv = 0 (see r_inv_of)
  DAB_ARG(h, cfg->world >= 1 && cfg->world <= DAB_MAX_RANKS && cfg->rank >= 0 && cfg->rank < cfg->world);
  DAB_ARG(h, cfg->system_dim % cfg->world == 0);
  DAB_ARG(h, cfg->n_obs_times >= 0 && cfg->n_obs >= 0);
  DAB_ARG(h, cfg->n_obs_times * cfg->n_obs == 0 ||
                 ((obs_values != nullptr || obs_values_dev != nullptr) && system_index != nullptr));
  DAB_ARG(h, cfg->n_cycles * cfg->t_pad == 0 || padded_time_idx != nullptr);
  DAB_ARG(h, cfg->world == 1 || cfg->n_steps <= l96_max_steps_per_launch());
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, cudaStreamSynchronize(h->stream));
  DAB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  Cycler& c = h->cyc;
  c.ready = false;
  cycler_close_peers(h);
  c.copied_pending[0] = c.copied_pending[1] = false;
  c.cfg = *cfg;
  const int k = cfg->ensemble_dim;
  c.n_loc = cfg->system_dim / cfg->world;
  c.n0 = c.n_loc * cfg->rank;
  const int maxs = l96_max_steps_per_launch();
  c.first_steps = cfg->n_steps < maxs ? cfg->n_steps : maxs;
  c.halo_l = 8 * c.first_steps;
  c.halo_r = 4 * c.first_steps;
  c.ld_e = (c.halo_l + c.n_loc + c.halo_r + 1) & ~1ll;
  // the reference's empty-R branch: no observation slot at all (T_pad * n_obs == 0)
  c.empty_R = ((long long)cfg->t_pad * cfg->n_obs == 0) || cfg->n_obs_times == 0;
  // grid-sharded runs are ordered by the statistics exchange alone (the halo reads of the transform
  // rely on it); without any observation slot there is no exchange, so such a run is refused
  if (cfg->world > 1 && c.empty_R)
    return fail(h, -1, "a grid-sharded run (world = %d) needs observations: the statistics exchange is its only inter-rank ordering", cfg->world);
  c.cur = 0;
  const size_t slab = sizeof(double) * (size_t)k * c.n_loc;
  if (slab > c.xb_bytes) {
    for (int w = 0; w < 2; ++w) {
      if (c.xb[w]) cudaFree(c.xb[w]);
      c.xb[w] = nullptr;
    }
    c.xb_bytes = 0;
    for (int w = 0; w < 2; ++w) DAB_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&c.xb[w]), slab));
    c.xb_bytes = slab;
  }
  for (int w = 0; w < 2; ++w) {
    DAB_CUDA(h, c.ext[w].reserve(sizeof(double) * (size_t)k * c.ld_e));
    if (!c.ev_xa[w]) DAB_CUDA(h, cudaEventCreateWithFlags(&c.ev_xa[w], cudaEventDisableTiming));
    if (!c.ev_copied[w]) DAB_CUDA(h, cudaEventCreateWithFlags(&c.ev_copied[w], cudaEventDisableTiming));
  }
  if (x0_host != nullptr)
    DAB_CUDA(h, cudaMemcpyAsync(c.xb[0], x0_host, slab, cudaMemcpyHostToDevice, h->stream));
  // ---- observations: CSR over observation times, rows owned by this rank, masked rows dropped
  const long long nt = cfg->n_obs_times, no = cfg->n_obs;
  c.csr_off.assign((size_t)nt + 1, 0);
  std::vector<int> loc;
  std::vector<double> val, swv;
  loc.reserve((size_t)(nt * no / cfg->world + 16));
  val.reserve((size_t)(nt * no / cfg->world + 16));
  c.has_sw = obs_sd != nullptr;
  // The statistics are sums over observation rows, so the rows of one observation time may be
  // visited in any order: order them by grid location (stable counting sort, O(n)) so that the 32
  // rows of a contraction block touch neighbouring columns of every member instead of 32 random
  // lines.  The (filtered, ordered) column list `src` is rebuilt only when the index row differs
  // from the previous time's (a stationary network builds it once); per time the values are then
  // one indexed copy.
  std::vector<int> src, lrow, tmp_o, tmp_l, count;
  std::vector<int> src_all;                 // device-resident values: the permutations, back to back ...
  std::vector<long long> src_base;          // ... and where the one of time t starts
  const bool dev_vals = obs_values_dev != nullptr;
  for (long long t = 0; t < nt; ++t) {
    const int64_t* irow = system_index + t * no;
    const double* vrow = dev_vals ? nullptr : obs_values + t * no;
    const bool reuse = t > 0 && cfg->stationary &&
                       memcmp(irow, irow - no, sizeof(int64_t) * (size_t)no) == 0;
    if (!reuse) {
      tmp_o.clear(); tmp_l.clear();
      bool sorted = true;
      for (long long o = 0; o < no; ++o) {
        if (!dev_vals && !cfg->stationary && std::isnan(vrow[o])) continue;   // loc mask, _dacycler.py:266-268
        const long long gi = irow[o];
        if (gi < 0 || gi >= cfg->system_dim)
          return fail(h, -1, "system_index[%lld][%lld] = %lld out of range", t, o, gi);
        if (gi < c.n0 || gi >= c.n0 + c.n_loc) continue;         // owned by another rank
        const int l = (int)(gi - c.n0);
        if (!tmp_l.empty() && l < tmp_l.back()) sorted = false;
        tmp_o.push_back((int)o); tmp_l.push_back(l);
      }
      const size_t nr = tmp_l.size();
      if (sorted || nr < 64) {
        src.swap(tmp_o); lrow.swap(tmp_l);
      } else {                 // bucket = location >> shift keeps the counting array small
        int shift = 0;
        while ((c.n_loc >> shift) > (long long)(4 * nr + 1024)) ++shift;
        const size_t nbk = (size_t)(c.n_loc >> shift) + 2;
        count.assign(nbk, 0);
        for (size_t i = 0; i < nr; ++i) count[(size_t)(tmp_l[i] >> shift) + 1]++;
        for (size_t bkt = 1; bkt < nbk; ++bkt) count[bkt] += count[bkt - 1];
        src.resize(nr); lrow.resize(nr);
        for (size_t i = 0; i < nr; ++i) {
          const size_t d = (size_t)count[(size_t)(tmp_l[i] >> shift)]++;
          src[d] = tmp_o[i]; lrow[d] = tmp_l[i];
        }
      }
    }
    const size_t nr = src.size(), base = loc.size();
    loc.resize(base + nr); val.resize(base + nr);
    memcpy(loc.data() + base, lrow.data(), sizeof(int) * nr);
    if (dev_vals) {
      if (!reuse) { src_base.push_back((long long)src_all.size()); src_all.insert(src_all.end(), src.begin(), src.end()); }
      else src_base.push_back(src_base.back());
    } else {
      for (size_t i = 0; i < nr; ++i) val[base + i] = vrow[src[i]];
    }
    if (obs_sd != nullptr) {
      swv.resize(base + nr);
      const double* srow = obs_sd + t * no;
      for (size_t i = 0; i < nr; ++i) { const double sg = srow[src[i]]; swv[base + i] = sg == 0.0 ? 0.0 : 1.0 / sg; }
    }
    c.csr_off[(size_t)t + 1] = (long long)loc.size();
  }
  // ---- compact Yb written by the forecast kernels: a stationary network whose observation times all have the same
  // rows (one permutation built, reused by every time), at most 2^31 rows, values not NaN-masked on the device
  c.emit_ok = false; c.yb_valid = false; c.n_compact = 0;
  bool want_emit = h->yb_emit;
  if (const char* ev = getenv("DAB_YB_EMIT")) want_emit = atoi(ev) != 0;        // read per set-up: tests switch it
  if (want_emit && cfg->stationary && nt > 0 && !src.empty() && c.n_loc < (1ll << 31)) {
    bool same = true;
    const long long nr0 = c.csr_off[1] - c.csr_off[0];
    for (long long t = 1; t < nt && same; ++t)
      same = (c.csr_off[(size_t)t + 1] - c.csr_off[(size_t)t]) == nr0 &&
             memcmp(system_index + t * no, system_index, sizeof(int64_t) * (size_t)no) == 0;
    bool ascending = true;
    for (size_t i = 1; i < (size_t)nr0 && ascending; ++i) ascending = loc[i] > loc[i - 1];      // strictly: one row per column
    int dgrid = 0, drc = 0;
    if (same && ascending && nr0 > 0 && contract_dense_plan(nr0, k, h->sms, &dgrid, &drc)) {
      const size_t nblk = (size_t)(c.n_loc >> 6) + 3;
      std::vector<int> rs(nblk, (int)nr0);
      size_t r = 0;
      for (size_t b = 0; b < nblk; ++b) {
        while (r < (size_t)nr0 && (long long)loc[r] < (long long)b * 64) ++r;
        rs[b] = (int)r;
      }
      c.emit_rows = nr0; c.emit_loc_off = 0; c.ldy = (nr0 + 1) & ~1ll;
      DAB_CUDA(h, c.row_start.reserve(sizeof(int) * nblk));
      DAB_CUDA(h, cudaMemcpyAsync(c.row_start.p, rs.data(), sizeof(int) * nblk, cudaMemcpyHostToDevice, h->stream));
      DAB_CUDA(h, c.yb.reserve(sizeof(double) * (size_t)k * c.ldy));
      DAB_CUDA(h, cudaStreamSynchronize(h->stream));            // rs goes out of scope
      c.emit_ok = true;
    }
  }
  DAB_CUDA(h, c.obs_loc.reserve(sizeof(int) * (loc.size() + 1)));
  DAB_CUDA(h, c.obs_val.reserve(sizeof(double) * (val.size() + 1)));
  if (!loc.empty()) {
    DAB_CUDA(h, cudaMemcpyAsync(c.obs_loc.p, loc.data(), sizeof(int) * loc.size(), cudaMemcpyHostToDevice, h->stream));
    if (!dev_vals) {
      DAB_CUDA(h, cudaMemcpyAsync(c.obs_val.p, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice, h->stream));
    } else {
      // the value table straight from the device-resident observations (and the NaN mask of a moving network)
      DAB_CUDA(h, h->s_src.reserve(sizeof(int) * (src_all.size() + 1)));
      DAB_CUDA(h, h->s_off.reserve(sizeof(long long) * 2 * ((size_t)nt + 1)));
      std::vector<long long> offs(2 * ((size_t)nt + 1), 0);
      for (long long t = 0; t <= nt; ++t) offs[(size_t)t] = c.csr_off[(size_t)t];
      for (long long t = 0; t < nt; ++t) offs[(size_t)nt + 1 + (size_t)t] = src_base[(size_t)t];
      DAB_CUDA(h, cudaMemcpyAsync(h->s_src.p, src_all.data(), sizeof(int) * src_all.size(), cudaMemcpyHostToDevice, h->stream));
      DAB_CUDA(h, cudaMemcpyAsync(h->s_off.p, offs.data(), sizeof(long long) * offs.size(), cudaMemcpyHostToDevice, h->stream));
      DAB_CUDA(h, pack_rows_launch(obs_values_dev, no, nt, h->s_src.as<int>(), h->s_off.as<long long>(),
                                   h->s_off.as<long long>() + nt + 1, cfg->stationary ? 0 : 1, c.obs_loc.as<int>(),
                                   c.obs_val.as<double>(), h->stream));
      DAB_CUDA(h, cudaStreamSynchronize(h->stream));           // offs / src_all go out of scope
    }
    if (c.has_sw) {
      DAB_CUDA(h, c.obs_sw.reserve(sizeof(double) * (swv.size() + 1)));
      DAB_CUDA(h, cudaMemcpyAsync(c.obs_sw.p, swv.data(), sizeof(double) * swv.size(), cudaMemcpyHostToDevice, h->stream));
    }
  }
  // ---- per-cycle segment tables from the padded time indices (0 = masked, t+1 otherwise)
  const int tp = cfg->t_pad > 0 ? cfg->t_pad : 1;
  std::vector<long long> seg((size_t)cfg->n_cycles * tp * 3 + 3, 0);
  c.n_seg.assign((size_t)cfg->n_cycles, 0);
  c.n_rows.assign((size_t)cfg->n_cycles, 0);
  c.seg_begin.assign((size_t)cfg->n_cycles, -1);
  long long max_rows = 0;
  for (int cy = 0; cy < cfg->n_cycles; ++cy) {
    long long rows = 0;
    int ns = 0;
    for (int s = 0; s < cfg->t_pad; ++s) {
      const long long ti = padded_time_idx[(long long)cy * cfg->t_pad + s];
      if (ti <= 0) continue;                                    // time mask, _dacycler.py:119
      if (ti > nt) return fail(h, -1, "padded_time_idx[%d][%d] = %lld exceeds n_obs_times", cy, s, ti);
      const long long b = c.csr_off[(size_t)ti - 1], e = c.csr_off[(size_t)ti];
      long long* sg = &seg[((size_t)cy * tp + ns) * 3];
      sg[0] = b; sg[1] = e - b; sg[2] = rows;
      rows += e - b;
      ++ns;
    }
    c.n_seg[(size_t)cy] = ns;
    c.n_rows[(size_t)cy] = rows;
    c.seg_begin[(size_t)cy] = ns == 1 ? seg[((size_t)cy * tp) * 3] : -1;
    if (rows > max_rows) max_rows = rows;
  }
  DAB_CUDA(h, c.seg.reserve(sizeof(long long) * seg.size()));
  DAB_CUDA(h, cudaMemcpyAsync(c.seg.p, seg.data(), sizeof(long long) * seg.size(), cudaMemcpyHostToDevice, h->stream));
  c.max_grid = contract_grid(max_rows, h->sms);
  DAB_CUDA(h, c.partial.reserve(sizeof(double) * contract_partial_doubles(k, c.max_grid)));
  DAB_CUDA(h, c.stats.reserve(sizeof(double) * ((size_t)k * k + k)));
  DAB_CUDA(h, cudaMemsetAsync(c.stats.p, 0, c.stats.bytes, h->stream));
  DAB_CUDA(h, c.T.reserve(sizeof(double) * (size_t)k * k));
  DAB_CUDA(h, c.lam.reserve(sizeof(double) * (size_t)k));
  DAB_CUDA(h, c.info.reserve(sizeof(int) * 4));
  if (cfg->world > 1) {
    DAB_CUDA(h, c.xchg.reserve(kXchgHeader + sizeof(double) * 2 * (size_t)cfg->world * ((size_t)k * k + k)));
    DAB_CUDA(h, cudaMemsetAsync(c.xchg.p, 0, c.xchg.bytes, h->stream));
    DAB_CUDA(h, c.xchg_ctr.reserve(sizeof(unsigned int) * 4));
    DAB_CUDA(h, cudaMemsetAsync(c.xchg_ctr.p, 0, c.xchg_ctr.bytes, h->stream));
  }
  c.xchg_seq = 0;
  c.truth = nullptr; c.sse_out = nullptr;
  DAB_CUDA(h, cudaStreamSynchronize(h->stream));               // host vectors go out of scope
  // ---- measure, don't guess: pick the forecast kernel's CTA width for this shape with CUDA events
  c.l96_T = 0;
  c.l96_chunks = 1;
  const long long tkey = (c.n_loc * 256 + k) * 64 + c.first_steps;
  for (size_t i = 0; i < h->tune_key.size(); ++i)
    if (h->tune_key[i] == tkey) { c.l96_T = h->tune_T[i]; c.l96_chunks = h->tune_chunks[i]; }
  if (c.l96_T == 0 && c.first_steps > 0 && (long long)k * c.n_loc >= (1ll << 17)) {
    DAB_CUDA(h, cudaMemsetAsync(c.ext[0].p, 0, c.ext[0].bytes, h->stream));
    cudaEvent_t e0, e1;
    DAB_CUDA(h, cudaEventCreate(&e0));
    DAB_CUDA(h, cudaEventCreate(&e1));
    const int n_cand = 20;
    // (the 6 x 64-thread build of the chained kernel, kL96V6Code + 1, was measured at every bench shape and never
    // won: it stays reachable through DAB_L96_V6=2)
    const int cand[n_cand] = {kL96V6Code, kL96V3Code + 32, kL96V3Code + 10, 128, 192, 256, 320, 384, 512,
                              l96_shape_code(16, 128), l96_shape_code(16, 160), l96_shape_code(16, 192),
                              l96_shape_code(16, 224), l96_shape_code(16, 256), l96_shape_code(17, 128),
                              l96_shape_code(17, 160), l96_shape_code(17, 192),
                              l96_shape_code(12, 128), l96_shape_code(12, 160), l96_shape_code(12, 192)};
    // splitting the window into 2 or 4 launches (forecast_ext, DAB_L96_CHUNKS) was measured at C3/C4 and
    // never won (shorter halos lose to the extra pass), so the tuner no longer tries it: that also keeps
    // the two extended scratch buffers it needs unallocated
    const int cand_chunks[3] = {1, 0, 0};
    float best_ms = 1e30f;
    for (int cc = 0; cc < 3; ++cc) {
      const int chunks = cand_chunks[cc];
      if (chunks < 1) continue;
      if (chunks > 1 && c.first_steps < 2 * chunks) continue;
      const int longest = (c.first_steps + chunks - 1) / chunks;
      for (int ci = 0; ci < n_cand; ++ci) {
        if (cand[ci] >= kL96V6Code) {
          if (chunks > 1 || !l96_v6_eligible(c.ext[0].as<double>(), c.xb[1], c.n_loc, c.ld_e, c.n_loc, c.halo_l, c.first_steps, false, cand[ci] - kL96V6Code)) continue;
        } else if (cand[ci] >= kL96V3Code) {
          if (chunks > 1 || !l96_v3_eligible(c.ext[0].as<double>(), c.xb[1], c.n_loc, c.ld_e, c.n_loc, c.halo_l, c.first_steps, false)) continue;
        } else {
          if (l96_shape_width(cand[ci]) < 12 * longest + 16) continue;
          if (chunks > 1 && cand[ci] >= 1000) continue;
        }
        float ms = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          DAB_CUDA(h, cudaEventRecord(e0, h->stream));
          DAB_CUDA(h, forecast_ext(h, c, c.ext[0].as<double>(), c.xb[1], nullptr, 0, c.first_steps, cand[ci],
                                   chunks, nullptr));
          DAB_CUDA(h, cudaEventRecord(e1, h->stream));
          DAB_CUDA(h, cudaEventSynchronize(e1));
          float t = 0.f;
          DAB_CUDA(h, cudaEventElapsedTime(&t, e0, e1));
          if (rep > 0 && t < ms) ms = t;
        }
        if (getenv("DAB_TUNE_VERBOSE")) fprintf(stderr, "[dab tune] n_loc=%lld k=%d steps=%d chunks=%d T=%d: %.1f us\n",
                                                c.n_loc, k, c.first_steps, chunks, cand[ci], ms * 1e3f);
        if (ms < best_ms) { best_ms = ms; c.l96_T = cand[ci]; c.l96_chunks = chunks; }
      }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->tune_key.push_back(tkey);
    h->tune_T.push_back(c.l96_T);
    h->tune_chunks.push_back(c.l96_chunks);
  }
  if (const char* ev = getenv("DAB_L96_V3")) {              // force (1..32 = steps per chunk) or forbid (0) the persistent kernel
    const int v = atoi(ev);
    if (v >= 1 && v <= 32) c.l96_T = kL96V3Code + v;
    else if (v == 0 && c.l96_T >= kL96V3Code) c.l96_T = 0;
  }
  if (const char* ev = getenv("DAB_L96_V6")) {              // force (1) or forbid (0) the chained kernel
    const int v = atoi(ev);
    if (v == 1 || v == 2) c.l96_T = kL96V6Code + v - 1;      // 1: 3 groups x 128 threads, 2: 6 groups x 64 threads
    else if (v == 0 && c.l96_T >= kL96V6Code) c.l96_T = 0;
  }
  if (const char* ev = getenv("DAB_L96_CHUNKS")) { const int v = atoi(ev); if (v >= 1 && v <= 8) c.l96_chunks = v; }
  h->launches = 0;
  c.ready = true;
  return 0;
}

int dab_cycler_setup(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values,
                     const int64_t* system_index, const int64_t* padded_time_idx) {
  return cycler_setup_impl(h, cfg, obs_values, system_index, padded_time_idx, nullptr);
}

int dab_cycler_setup_dev(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values_dev,
                         const int64_t* system_index, const int64_t* padded_time_idx) {
  DAB_ARG(h, h != nullptr && cfg != nullptr);
  DAB_ARG(h, cfg->n_obs_times * cfg->n_obs == 0 || obs_values_dev != nullptr);
  return cycler_setup_impl(h, cfg, nullptr, system_index, padded_time_idx, nullptr, nullptr, obs_values_dev);
}

int dab_cycler_setup_diag(dab_handle* h, const dab_cycle_config* cfg, const double* obs_values,
                          const int64_t* system_index, const int64_t* padded_time_idx, const double* obs_error_sd) {
  return cycler_setup_impl(h, cfg, obs_values, system_index, padded_time_idx, nullptr, obs_error_sd);
}

int dab_cycler_set_state(dab_handle* h, const double* x, int is_host) {
  DAB_ARG(h, h != nullptr && h->cyc.ready && x != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  Cycler& c = h->cyc;
  c.yb_valid = false;                      // the compact copy belonged to the background this call replaces
  DAB_CUDA(h, cudaMemcpyAsync(c.xb[c.cur], x, sizeof(double) * (size_t)c.cfg.ensemble_dim * c.n_loc,
                              is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

int dab_cycler_get_state(dab_handle* h, double* x, int is_host) {
  DAB_ARG(h, h != nullptr && h->cyc.ready && x != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  Cycler& c = h->cyc;
  DAB_CUDA(h, cudaMemcpyAsync(x, c.xb[c.cur], sizeof(double) * (size_t)c.cfg.ensemble_dim * c.n_loc,
                              is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
  if (is_host) DAB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int dab_cycler_stats(dab_handle* h, int cycle) {
  DAB_ARG(h, h != nullptr && h->cyc.ready);
  Cycler& c = h->cyc;
  DAB_ARG(h, cycle >= 0 && cycle < c.cfg.n_cycles);
  DAB_CUDA(h, cudaSetDevice(h->device));
  if (c.empty_R) return 0;
  const int k = c.cfg.ensemble_dim;
  const int tp = c.cfg.t_pad > 0 ? c.cfg.t_pad : 1;
  const long long rows = c.n_rows[(size_t)cycle];
  const int grid = contract_grid(rows, h->sms);
  const int ns = c.n_seg[(size_t)cycle] > 0 ? c.n_seg[(size_t)cycle] : 1;
  prof_mark(h, PROF_START);
  int dgrid = 0, drc = 0;
  const long long seg_begin = c.seg_begin[(size_t)cycle];
  if (c.yb_valid && c.n_seg[(size_t)cycle] == 1 && rows == c.emit_rows && seg_begin >= 0 &&
      contract_dense_plan(rows, k, h->sms, &dgrid, &drc) && dgrid == grid) {
    // the forecast that made this background also wrote its observed columns, compact: no gather
    ++c.n_compact;
    DAB_CUDA(h, contract_dense_launch(c.yb.as<double>(), c.ldy, k, c.obs_val.as<double>() + seg_begin,
                                      c.has_sw ? c.obs_sw.as<double>() + seg_begin : nullptr, rows, grid, drc,
                                      c.partial.as<double>(), true, h->stream));
  } else {
    DAB_CUDA(h, contract_launch(c.xb[c.cur], c.n_loc, k, c.obs_loc.as<int>(), c.obs_val.as<double>(),
                                c.has_sw ? c.obs_sw.as<double>() : nullptr,
                                c.seg.as<long long>() + (size_t)cycle * tp * 3, ns, rows,
                                c.partial.as<double>(), grid, true, h->stream));
  }
  prof_mark(h, PROF_CONTRACT);
  const double r_inv = c.has_sw ? 1.0 : r_inv_of(c.cfg.obs_error_sd);
  if (stats_exchange_ready(c)) {
    // reduction fused with the all-gather of the ranks' statistics over peer memory, then the
    // rank-ordered sum (replaces reduce + NCCL all-reduce; see stats_reduce_kernel)
    const int G = c.cfg.world, me = c.cfg.rank;
    const size_t len = (size_t)k * k + k;
    const unsigned long long seq = ++c.xchg_seq;
    const int par = (int)(seq & 1);
    StatsPush push{};
    push.world = G; push.seq = seq; push.done_ctr = c.xchg_ctr.as<unsigned int>();
    for (int g = 0; g < G; ++g) {
      char* base = g == me ? c.xchg.as<char>() : reinterpret_cast<char*>(const_cast<double*>(c.peer[g][2]));
      push.flag[g] = reinterpret_cast<unsigned long long*>(base) + par * kMaxRanks + me;
      push.slot[g] = reinterpret_cast<double*>(base + kXchgHeader) + ((size_t)par * G + me) * len;
    }
    DAB_CUDA(h, stats_reduce_launch(c.partial.as<double>(), grid, k, r_inv, nullptr, &push, h->stream));
    DAB_CUDA(h, stats_gather_launch(c.xchg.as<unsigned long long>() + par * kMaxRanks, seq,
                                    reinterpret_cast<const double*>(c.xchg.as<char>() + kXchgHeader) + (size_t)par * G * len,
                                    k, G, c.stats.as<double>(), h->xchg_timeout_ns, h->stream));
    h->launches += 1;
  } else {
    DAB_CUDA(h, stats_reduce_launch(c.partial.as<double>(), grid, k, r_inv, c.stats.as<double>(), nullptr, h->stream));
  }
  prof_mark(h, PROF_REDUCE);
  h->launches += 2;
  return 0;
}

int dab_cycler_stats_exchange(dab_handle* h) {
  return (h && h->cyc.ready && stats_exchange_ready(h->cyc)) ? 1 : 0;
}

double* dab_cycler_stats_ptr(dab_handle* h) { return (h && h->cyc.ready) ? h->cyc.stats.as<double>() : nullptr; }
void* dab_cycler_stream(dab_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }
double* dab_cycler_background_ptr(dab_handle* h) { return (h && h->cyc.ready) ? h->cyc.xb[h->cyc.cur] : nullptr; }
int dab_cycler_kernel_launches(dab_handle* h) { return h ? (int)h->launches : 0; }
int dab_cycler_compact_contractions(dab_handle* h) { return (h && h->cyc.ready) ? h->cyc.n_compact : 0; }
int dab_cycler_forecast_shape(dab_handle* h) { return (h && h->cyc.ready) ? h->cyc.l96_T * 10 + h->cyc.l96_chunks : 0; }

// eig (K4) + transform (K5, fused halo exchange) + forecast (K1).  The analysis of this cycle is
// left in ext[cycle & 1]; out_host / out_dev optionally receive its slab part.
static int cycler_update(dab_handle* h, int cycle, double* out_host, double* out_dev, double* out_traj) {
  Cycler& c = h->cyc;
  const int k = c.cfg.ensemble_dim;
  const int eb = cycle & 1;
  prof_mark(h, PROF_START);
  DAB_CUDA(h, eig_transform_launch(c.stats.as<double>(), k, c.cfg.rho, c.empty_R, c.T.as<double>(),
                                   nullptr, c.info.as<int>(), h->ns_ws.as<double>(), h->eig_method, h->stream));
  prof_mark(h, PROF_EIG);
  if (k > 64 && !c.empty_R && h->eig_method != kEigJacobi) h->launches += 1;   // Newton-Schulz + the (flag-gated) Jacobi fallback kernel
  if (c.copied_pending[eb]) {                 // the D2H of cycle-2's analysis still reads ext[eb]
    DAB_CUDA(h, cudaStreamWaitEvent(h->stream, c.ev_copied[eb], 0));
    c.copied_pending[eb] = false;
  }
  TransformArgs a{};
  a.T = c.T.as<double>();
  for (int r = 0; r < c.cfg.world; ++r) {
    if (r == c.cfg.rank) a.src[r] = c.xb[c.cur];
    else {
      if (c.peer[r][c.cur] == nullptr)
        return fail(h, -3, "peer buffer of rank %d not registered (dab_cycler_ipc_import / dab_cycler_set_peer)", r);
      a.src[r] = c.peer[r][c.cur];
    }
  }
  a.src_ld = c.n_loc;
  a.dst = c.ext[eb].as<double>(); a.dst_ld = c.ld_e; a.dst_base = c.halo_l;
  a.n_loc = c.n_loc; a.n_global = c.cfg.system_dim; a.n0 = c.n0;
  a.halo_l = c.halo_l; a.halo_r = c.halo_r; a.k = k; a.rank = c.cfg.rank;
  DAB_CUDA(h, transform_launch(a, h->sms, h->stream));
  prof_mark(h, PROF_TRANSFORM);
  h->launches += 2;
  const double* xa = c.ext[eb].as<double>() + c.halo_l;
  if (c.truth != nullptr && c.sse_out != nullptr) {          // score this analysis on the device
    DAB_CUDA(h, mean_sqerr_launch(xa, c.ld_e, k, c.n_loc, c.truth + (size_t)cycle * c.n_loc,
                                  c.sse_partial.as<double>(), c.sse_out + cycle, h->stream));
    h->launches += 2;
  }
  if (out_host) {
    DAB_CUDA(h, cudaEventRecord(c.ev_xa[eb], h->stream));
    DAB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, c.ev_xa[eb], 0));
    DAB_CUDA(h, cudaMemcpy2DAsync(out_host, sizeof(double) * c.n_loc, xa, sizeof(double) * c.ld_e,
                                  sizeof(double) * c.n_loc, k, cudaMemcpyDeviceToHost, h->copy_stream));
    DAB_CUDA(h, cudaEventRecord(c.ev_copied[eb], h->copy_stream));
    c.copied_pending[eb] = true;
  }
  if (out_dev) {
    DAB_CUDA(h, cudaMemcpy2DAsync(out_dev, sizeof(double) * c.n_loc, xa, sizeof(double) * c.ld_e,
                                  sizeof(double) * c.n_loc, k, cudaMemcpyDeviceToDevice, h->stream));
  }
  // forecast: ext[eb] -> xb[cur ^ 1]
  prof_mark(h, PROF_START);
  const int S = c.cfg.n_steps;
  double* nxt = c.xb[c.cur ^ 1];
  const long long traj_stride = (long long)k * c.n_loc;
  if (S == 0 || S > c.first_steps) c.yb_valid = false;
  if (S == 0) {
    DAB_CUDA(h, cudaMemcpy2DAsync(nxt, sizeof(double) * c.n_loc, xa, sizeof(double) * c.ld_e,
                                  sizeof(double) * c.n_loc, k, cudaMemcpyDeviceToDevice, h->stream));
  } else if (S <= c.first_steps) {
    int nl = 0;
    ObsEmit emit;
    bool emitted = false;
    if (c.emit_ok) {
      emit.loc = c.obs_loc.as<int>() + c.emit_loc_off; emit.row_start = c.row_start.as<int>();
      emit.yb = c.yb.as<double>(); emit.ldy = c.ldy; emit.n_rows = (int)c.emit_rows;
    }
    c.yb_valid = false;
    DAB_CUDA(h, forecast_ext(h, c, c.ext[eb].as<double>(), nxt, out_traj, traj_stride, S, c.l96_T, c.l96_chunks, &nl,
                             c.emit_ok ? &emit : nullptr, &emitted));
    c.yb_valid = emitted;
    h->launches += nl;
  } else {
    // long window on one rank: first chunk from the extended buffer, the rest periodic.
    // chunk outputs alternate between the two Xb buffers and a scratch so the last lands in nxt.
    DAB_CUDA(h, h->s_partial.reserve(sizeof(double) * (size_t)k * c.n_loc));
    double* tmp = h->s_partial.as<double>();
    const int maxs = l96_max_steps_per_launch();
    const int rest = S - c.first_steps;
    const int rest_chunks = (rest + maxs - 1) / maxs;
    double* first_dst = (rest_chunks % 2 == 0) ? nxt : tmp;
    DAB_CUDA(h, l96_forecast_launch(c.ext[eb].as<double>(), first_dst, out_traj, c.n_loc, k, c.first_steps,
                                    c.cfg.model_dt, c.cfg.forcing, c.ld_e, c.n_loc, false, c.halo_l,
                                    -(long long)c.halo_l, c.n_loc + c.halo_r, traj_stride, h->sms,
                                    c.l96_T >= kL96V3Code ? 0 : c.l96_T, h->stream));
    h->launches++;
    const double* src = first_dst;
    int done = c.first_steps;
    for (int ch = 0; ch < rest_chunks; ++ch) {
      const int s = (S - done) < maxs ? (S - done) : maxs;
      double* dst = (src == tmp) ? nxt : tmp;
      DAB_CUDA(h, l96_forecast_launch(src, dst, out_traj ? out_traj + (long long)done * traj_stride : nullptr,
                                      c.n_loc, k, s, c.cfg.model_dt, c.cfg.forcing, c.n_loc, c.n_loc, true,
                                      0, 0, 0, traj_stride, h->sms, 0, h->stream));
      h->launches++;
      src = dst;
      done += s;
    }
  }
  prof_mark(h, PROF_FORECAST);
  c.cur ^= 1;
  return 0;
}

int dab_cycler_update(dab_handle* h, int cycle, double* out_analysis_dev) {
  DAB_ARG(h, h != nullptr && h->cyc.ready);
  DAB_ARG(h, cycle >= 0 && cycle < h->cyc.cfg.n_cycles);
  DAB_CUDA(h, cudaSetDevice(h->device));
  return cycler_update(h, cycle, nullptr, out_analysis_dev, nullptr);
}

int dab_cycler_set_truth(dab_handle* h, const double* truth_dev, double* sse_out_dev) {
  DAB_ARG(h, h != nullptr && h->cyc.ready);
  DAB_ARG(h, (truth_dev == nullptr) == (sse_out_dev == nullptr));
  DAB_CUDA(h, cudaSetDevice(h->device));
  if (truth_dev != nullptr) DAB_CUDA(h, h->cyc.sse_partial.reserve(sizeof(double) * 512));
  h->cyc.truth = truth_dev;
  h->cyc.sse_out = sse_out_dev;
  return 0;
}

int dab_cycler_update_host(dab_handle* h, int cycle, double* out_analysis_host) {
  DAB_ARG(h, h != nullptr && h->cyc.ready);
  DAB_ARG(h, cycle >= 0 && cycle < h->cyc.cfg.n_cycles);
  DAB_CUDA(h, cudaSetDevice(h->device));
  return cycler_update(h, cycle, out_analysis_host, nullptr, nullptr);
}

int dab_cycler_run(dab_handle* h, int c0, int c1, double* out_analysis, double* out_traj) {
  DAB_ARG(h, h != nullptr && h->cyc.ready);
  Cycler& c = h->cyc;
  DAB_ARG(h, c.cfg.world == 1);
  DAB_ARG(h, c0 >= 0 && c1 >= c0 && c1 <= c.cfg.n_cycles);
  DAB_CUDA(h, cudaSetDevice(h->device));
  const size_t slab = (size_t)c.cfg.ensemble_dim * c.n_loc;
  for (int cy = c0; cy < c1; ++cy) {
    int rc = dab_cycler_stats(h, cy);
    if (rc) return rc;
    rc = cycler_update(h, cy, out_analysis ? out_analysis + (size_t)(cy - c0) * slab : nullptr, nullptr,
                       out_traj ? out_traj + (size_t)(cy - c0) * c.cfg.n_steps * slab : nullptr);
    if (rc) return rc;
  }
  return 0;
}

int dab_cycler_ipc_export(dab_handle* h, int which, void* handle_out) {
  DAB_ARG(h, h != nullptr && h->cyc.ready && handle_out != nullptr && which >= 0 && which <= 2);
  DAB_ARG(h, which < 2 || h->cyc.xchg.p != nullptr);
  static_assert(sizeof(cudaIpcMemHandle_t) == DAB_IPC_HANDLE_BYTES, "IPC handle size");
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  DAB_CUDA(h, cudaIpcGetMemHandle(&mh, which < 2 ? static_cast<void*>(h->cyc.xb[which]) : h->cyc.xchg.p));
  memcpy(handle_out, &mh, sizeof(mh));
  return 0;
}

int dab_cycler_ipc_import(dab_handle* h, int peer_rank, int which, const void* handle_in) {
  DAB_ARG(h, h != nullptr && h->cyc.ready && handle_in != nullptr && which >= 0 && which <= 2);
  DAB_ARG(h, peer_rank >= 0 && peer_rank < h->cyc.cfg.world && peer_rank != h->cyc.cfg.rank);
  DAB_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle_in, sizeof(mh));
  void* p = nullptr;
  DAB_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
  h->cyc.peer[peer_rank][which] = static_cast<const double*>(p);
  h->cyc.peer_open[peer_rank][which] = true;
  return 0;
}

double* dab_cycler_buffer_ptr(dab_handle* h, int which) {
  if (!h || !h->cyc.ready || which < 0 || which > 2) return nullptr;
  return which < 2 ? h->cyc.xb[which] : h->cyc.xchg.as<double>();
}

int dab_cycler_set_peer(dab_handle* h, int peer_rank, int which, const double* ptr) {
  DAB_ARG(h, h != nullptr && h->cyc.ready && ptr != nullptr && which >= 0 && which <= 2);
  DAB_ARG(h, peer_rank >= 0 && peer_rank < h->cyc.cfg.world && peer_rank != h->cyc.cfg.rank);
  Cycler& c = h->cyc;
  if (c.peer_open[peer_rank][which]) return fail(h, -1, "peer %d buffer %d already imported through IPC", peer_rank, which);
  c.peer[peer_rank][which] = ptr;
  return 0;
}

int dab_etkf_cycle_batched_f64(dab_handle* h, const dab_cycle_config* cfg, int batch, const double* rho,
                               const double* x0, const double* obs_values, const int64_t* system_index,
                               int shared_index, const int64_t* padded_time_idx, double* out_analysis,
                               double* out_mean, double* out_final) {
  DAB_ARG(h, h != nullptr && cfg != nullptr && batch >= 1 && rho != nullptr && x0 != nullptr);
  DAB_ARG(h, cfg->world == 1 && cfg->rank == 0);
  DAB_ARG(h, cfg->system_dim >= 4 && cfg->ensemble_dim >= 2 && cfg->ensemble_dim <= 32);
  DAB_ARG(h, cfg->system_dim * cfg->ensemble_dim <= 4096);
  DAB_ARG(h, cfg->n_steps >= 0 && cfg->n_cycles >= 0 && cfg->t_pad >= 0 && cfg->n_obs >= 0);
  DAB_ARG(h, (long long)cfg->t_pad * cfg->n_obs * cfg->ensemble_dim <= 8192);
  DAB_ARG(h, cfg->model_dt > 0.0 && cfg->obs_error_sd == cfg->obs_error_sd);
  DAB_ARG(h, cfg->n_obs_times * cfg->n_obs == 0 || (obs_values != nullptr && system_index != nullptr));
  DAB_ARG(h, cfg->n_cycles * cfg->t_pad == 0 || padded_time_idx != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  const int k = cfg->ensemble_dim, N = (int)cfg->system_dim;
  const size_t kN = (size_t)k * N;
  const size_t n_obs_all = (size_t)cfg->n_obs_times * cfg->n_obs;
  const bool shared_vals = (shared_index & 2) != 0;       // bit 1: one observation set for all experiments
  shared_index &= 1;
  for (int b = 0; b < batch; ++b) DAB_ARG(h, rho[b] > 0.0);
  for (size_t i = 0; i < n_obs_all * (shared_index ? 1 : (size_t)batch); ++i)
    if (system_index[i] < 0 || system_index[i] >= N) return fail(h, -1, "system_index[%zu] = %lld out of range", i, (long long)system_index[i]);
  for (long long i = 0; i < (long long)cfg->n_cycles * cfg->t_pad; ++i)
    if (padded_time_idx[i] < 0 || padded_time_idx[i] > cfg->n_obs_times)
      return fail(h, -1, "padded_time_idx[%lld] = %lld out of range", i, (long long)padded_time_idx[i]);
  const size_t smem = batched_smem_bytes(k, N, cfg->t_pad, (int)cfg->n_obs);
  if (smem > 227 * 1024) return fail(h, -1, "batched path: experiment needs %zu bytes of shared memory (> 227 KB)", smem);
  cudaStream_t st = h->stream;
  auto up = [&](DevBuf& buf, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = buf.reserve(bytes ? bytes : 8);
    if (e != cudaSuccess || bytes == 0) return e;
    return cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st);
  };
  DAB_CUDA(h, up(h->b_x0, x0, sizeof(double) * kN * batch));
  DAB_CUDA(h, up(h->b_val, obs_values, sizeof(double) * n_obs_all * (shared_vals ? 1 : (size_t)batch)));
  DAB_CUDA(h, up(h->b_idx, system_index, sizeof(int64_t) * n_obs_all * (shared_index ? 1 : batch)));
  DAB_CUDA(h, up(h->b_pad, padded_time_idx, sizeof(int64_t) * (size_t)cfg->n_cycles * cfg->t_pad));
  DAB_CUDA(h, up(h->b_rho, rho, sizeof(double) * batch));
  BatchedArgs a{};
  a.x0 = h->b_x0.as<double>(); a.obs_val = h->b_val.as<double>();
  a.obs_idx = h->b_idx.as<long long>(); a.idx_batch_stride = shared_index ? 0 : (long long)n_obs_all;
  a.val_batch_stride = shared_vals ? 0 : (long long)n_obs_all;
  a.padded = h->b_pad.as<long long>(); a.rho = h->b_rho.as<double>();
  if (out_analysis) { DAB_CUDA(h, h->b_out.reserve(sizeof(double) * kN * batch * cfg->n_cycles)); a.out_analysis = h->b_out.as<double>(); }
  if (out_mean) { DAB_CUDA(h, h->b_mean.reserve(sizeof(double) * (size_t)N * batch * cfg->n_cycles)); a.out_mean = h->b_mean.as<double>(); }
  if (out_final) { DAB_CUDA(h, h->b_fin.reserve(sizeof(double) * kN * batch)); a.out_final = h->b_fin.as<double>(); }
  a.n_obs_times = cfg->n_obs_times;
  a.k = k; a.N = N; a.n_steps = cfg->n_steps; a.n_cycles = cfg->n_cycles; a.t_pad = cfg->t_pad;
  a.n_obs = (int)cfg->n_obs;
  a.empty_R = ((long long)cfg->t_pad * cfg->n_obs == 0) || cfg->n_obs_times == 0;
  a.force_jacobi = h->eig_method == kEigJacobi;
  a.dt = cfg->model_dt; a.F = cfg->forcing; a.r_inv = r_inv_of(cfg->obs_error_sd); a.stationary = cfg->stationary;
  DAB_CUDA(h, batched_launch(a, batch, st));
  h->launches++;
  if (out_analysis) DAB_CUDA(h, cudaMemcpyAsync(out_analysis, a.out_analysis, sizeof(double) * kN * batch * cfg->n_cycles, cudaMemcpyDeviceToHost, st));
  if (out_mean) DAB_CUDA(h, cudaMemcpyAsync(out_mean, a.out_mean, sizeof(double) * (size_t)N * batch * cfg->n_cycles, cudaMemcpyDeviceToHost, st));
  if (out_final) DAB_CUDA(h, cudaMemcpyAsync(out_final, a.out_final, sizeof(double) * kN * batch, cudaMemcpyDeviceToHost, st));
  DAB_CUDA(h, cudaStreamSynchronize(st));
  return 0;
}

int dab_cycler_set_profiling(dab_handle* h, int on) {
  DAB_ARG(h, h != nullptr);
  h->profiling = on != 0;
  h->prof_used = 0;
  return 0;
}

int dab_cycler_read_profile(dab_handle* h, double* ms_out, int* count_out) {
  DAB_ARG(h, h != nullptr && ms_out != nullptr && count_out != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 5; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
  for (size_t i = 1; i < h->prof_used; ++i) {
    const int tag = h->prof_tags[i];
    if (tag < 0) continue;
    float ms = 0.f;
    DAB_CUDA(h, cudaEventElapsedTime(&ms, h->prof_pool[i - 1], h->prof_pool[i]));
    ms_out[tag] += ms;
    count_out[tag] += 1;
  }
  h->prof_used = 0;
  return 0;
}

int dab_measure_fp64_peak(dab_handle* h, double* dfma_tflops, double* dmma_tflops) {
  DAB_ARG(h, h != nullptr && dfma_tflops != nullptr && dmma_tflops != nullptr);
  DAB_CUDA(h, cudaSetDevice(h->device));
  DAB_CUDA(h, fp64_peak_measure(h->sms, dfma_tflops, dmma_tflops, h->stream));
  return 0;
}

static int cycle_host_impl(dab_handle* h, const dab_cycle_config* cfg, const double* x0,
                           const double* obs_values, const int64_t* system_index,
                           const int64_t* padded_time_idx, const double* obs_sd, double* out_analysis) {
  DAB_ARG(h, h != nullptr && cfg != nullptr && x0 != nullptr);
  DAB_ARG(h, cfg->world == 1);
  int rc = cycler_setup_impl(h, cfg, obs_values, system_index, padded_time_idx, x0, obs_sd);   // uploads x0 too
  if (rc) return rc;
  // pin the output range if the caller did not, so the per-cycle D2H overlaps the next cycle
  bool registered = false;
  const size_t out_bytes = sizeof(double) * (size_t)cfg->n_cycles * cfg->ensemble_dim * cfg->system_dim;
  if (out_analysis && out_bytes) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, out_analysis);
    if (e != cudaSuccess) cudaGetLastError();
    if (e != cudaSuccess || at.type == cudaMemoryTypeUnregistered) {
      if (cudaHostRegister(out_analysis, out_bytes, cudaHostRegisterDefault) == cudaSuccess) registered = true;
      else cudaGetLastError();
    }
  }
  rc = dab_cycler_run(h, 0, cfg->n_cycles, out_analysis, nullptr);
  if (!rc) rc = dab_sync(h);
  if (registered) cudaHostUnregister(out_analysis);
  return rc;
}

int dab_etkf_cycle_host_f64(dab_handle* h, const dab_cycle_config* cfg, const double* x0,
                            const double* obs_values, const int64_t* system_index,
                            const int64_t* padded_time_idx, double* out_analysis) {
  return cycle_host_impl(h, cfg, x0, obs_values, system_index, padded_time_idx, nullptr, out_analysis);
}

int dab_etkf_cycle_host_diag_f64(dab_handle* h, const dab_cycle_config* cfg, const double* x0,
                                 const double* obs_values, const int64_t* system_index,
                                 const int64_t* padded_time_idx, const double* obs_error_sd, double* out_analysis) {
  return cycle_host_impl(h, cfg, x0, obs_values, system_index, padded_time_idx, obs_error_sd, out_analysis);
}

}  // extern "C"
