// Shared helpers for the B200 (sm_100a) ETKF kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace dab {

constexpr int kMaxEns = 128;          // k <= 128 (north star: eigensolver held in one CTA)

// FP64 tensor-core MMA, the native sm_100a shape (SASS: DMMA.8x8x4).
//   A 8x4 row-major : lane l holds A[l/4][l%4]
//   B 4x8 col-major : lane l holds B[l%4][l/4]
//   C/D 8x8         : lane l holds C[l/4][2*(l%4)], C[l/4][2*(l%4)+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ long long pmod(long long a, long long n) {
  long long r = a % n;
  return r < 0 ? r + n : r;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------
// The five kernels of a cycle run back to back on one stream and each is short (8-130 us), so
// the drain -> launch -> ramp gap at every boundary is a visible share of the cycle.  A kernel
// launched through `launch_chain` may become resident while its predecessor is still running;
// it does whatever does not depend on the predecessor (index tables, barrier set-up), then
// `pdl_wait()` blocks until the predecessor has completed and its writes are visible.  Every
// kernel calls `pdl_launch_dependents()` only AFTER its own wait, so that at most two kernels
// share the machine and "my predecessor finished" implies "everything before it finished":
// reads and writes after the wait keep plain stream-order semantics.  Both instructions are
// no-ops in a kernel launched without the attribute (DAB_PDL=0, or an ordinary <<<>>> launch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // engine.cu: DAB_PDL != "0"

// Forecast kernels: write the observed columns of a finished tile into the compact Yb row of `member` (kernels.h:
// ObsEmit).  The tile's values sit in shared memory, element e (window-relative) at tile[pad(e)]; [p, q) is the range
// of slab columns the tile stores; g0 the slab column of window position 0.  Every thread of the group / CTA calls it
// (tid of nthreads).  The candidate rows are those of the 64-column blocks that meet [p, q): no dependent search.
template <typename Pad>
__device__ __forceinline__ void emit_observed(const int* __restrict__ loc, const int* __restrict__ row_start, int n_rows,
                                              double* __restrict__ yb_row, const double* tile, Pad pad, long long g0,
                                              long long p, long long q, int tid, int nthreads) {
  if (q <= p) return;
  const int r_lo = row_start[p >> 6], r_hi = row_start[((q - 1) >> 6) + 1];
  for (int r = r_lo + tid; r < r_hi && r < n_rows; r += nthreads) {
    const long long c = loc[r];
    if (c >= p && c < q) yb_row[r] = tile[pad((int)(c - g0))];
  }
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(bool chain, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (chain && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace dab
