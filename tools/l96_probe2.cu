// Round-2 forecast probe: what bounds the FP64 pipe for the 17-instruction RK4 stage arithmetic at
// P = 16 points per thread, and what the warp-private window (shuffles per stage, one shared-memory
// halo refresh per STEP) costs on top.  Not part of the product; results in profiles/.
//   CM 0: coefficients live in registers (derived from dt in the kernel, as the round-1 kernel)
//   CM 1: coefficients are kernel parameters (constant bank / uniform registers: a DFMA then reads
//         two register operands instead of three)
//   XM 0: neighbours from the thread's own registers (arithmetic only)
//   XM 1: warp shuffles per stage, nothing else
//   XM 2: shuffles + per-step halo refresh through shared memory, one __syncthreads per step
//   XM 3: shuffles + per-step halo refresh, neighbour-only mbarriers (no CTA barrier)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct Coef { double h2, h3, h6, dt, h2F, dtF; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_test_wait(unsigned long long* b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void flag_wait(const volatile int* f, int v) { while (*f < v) { } }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{ .reg .pred p;\n"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W; }" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int CM, int XM, int MAXT, int MINB, int NCH = 1, int P = 16>
__global__ void __launch_bounds__(MAXT, MINB) probe(double* out, int n_steps, double dt_in, double F_in, const Coef c, int stagger) {
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, NW = blockDim.x >> 5;
  // per warp and parity: 8 doubles for the right neighbour's left halo, 4 for the left neighbour's right halo
  double* edge = smem;                                           // [2][NW][12]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 2 * NW * 12);   // [2][NW]
  double s[P], xh[P], xd[P], acc[P];
  double h2, h3, h6, dt, h2F, dtF;
  if (CM == 0) { dt = dt_in; h2 = 0.5 * dt; h6 = dt / 6.0; h3 = dt / 3.0; h2F = h2 * F_in; dtF = dt * F_in; }
  else { dt = c.dt; h2 = c.h2; h6 = c.h6; h3 = c.h3; h2F = c.h2F; dtF = c.dtF; }
#pragma unroll
  for (int i = 0; i < P; ++i) { s[i] = 8.0 + 0.01 * ((t * P + i) % 37) + blockIdx.x * 1e-6; xh[i] = s[i] + h2F; xd[i] = s[i] + dtF; }
  if (XM == 3 || XM == 4) {
    if (t < 2 * NW) mbar_init(&bars[t], 1);
    __syncthreads();
  }
  if (XM == 5) {
    if (t < NW) reinterpret_cast<volatile int*>(bars)[t] = 0;
    __syncthreads();
  }
  // chains: NCH independent groups of consecutive warps inside the CTA (own named barrier, no exchange
  // across the group boundary), started `stagger` cycles apart so that their per-step sync bubbles do
  // not coincide; with NCH == 1 the stagger goes by blockIdx (co-resident CTAs)
  const int wpc = NW / NCH, chain = w / wpc, wc = w - chain * wpc;
  if (stagger > 0) {
    const long long t0 = clock64();
    const long long d = (long long)stagger * (NCH > 1 ? chain : (int)(blockIdx.x % MINB));
    while (clock64() - t0 < d) { }
  }
  for (int step = 0; step < n_steps; ++step) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      double Lx, Ly, R;
      if (XM >= 1) {
        Lx = __shfl_up_sync(0xffffffffu, s[P - 2], 1); Ly = __shfl_up_sync(0xffffffffu, s[P - 1], 1);
        R = __shfl_down_sync(0xffffffffu, s[0], 1);
      } else { Lx = s[P - 2]; Ly = s[P - 1]; R = s[0]; }
      double kk[P];
      kk[0] = fma(s[1] - Lx, Ly, -s[0]);
      kk[1] = fma(s[2] - Ly, s[0], -s[1]);
      kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
#pragma unroll
      for (int i = 2; i < P - 1; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (CM == 0) {
          const double ca = (st == 2) ? dt : h2, cw = (st == 0 || st == 3) ? h6 : h3;
          if (st == 0) acc[i] = fma(cw, kk[i], xd[i]);
          else if (st < 3) acc[i] = fma(cw, kk[i], acc[i]);
          if (st < 3) s[i] = fma(ca, kk[i], st == 2 ? xd[i] : xh[i]);
          else s[i] = fma(cw, kk[i], acc[i]);
        } else {
          if (st == 0) { acc[i] = fma(kk[i], c.h6, xd[i]); s[i] = fma(kk[i], c.h2, xh[i]); }
          else if (st == 1) { acc[i] = fma(kk[i], c.h3, acc[i]); s[i] = fma(kk[i], c.h2, xh[i]); }
          else if (st == 2) { acc[i] = fma(kk[i], c.h3, acc[i]); s[i] = fma(kk[i], c.dt, xd[i]); }
          else s[i] = fma(kk[i], c.h6, acc[i]);
        }
      }
    }
    if (XM >= 2) {
      const int b = step & 1;
      double* mine = edge + (b * NW + w) * 12;
      if (lane == 31) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(mine + i) = make_double2(s[P - 12 + i], s[P - 11 + i]);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; i += 2) *reinterpret_cast<double2*>(mine + 8 + i) = make_double2(s[8 + i], s[9 + i]);
      }
      if (XM == 2) __syncthreads();
      else if (XM == 6) asm volatile("bar.sync %0, %1;" ::"r"(chain + 1), "r"(wpc * 32) : "memory");
      else if (XM == 5) {                       // monotonic step counters in shared memory; the WHOLE warp polls
        unsigned* flags = reinterpret_cast<unsigned*>(bars);   // (a spin loop under `if (lane == ..)` leaves the warp split)
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(flags + w)), "r"(step + 1) : "memory");
        const unsigned fl = smem_u32(flags + (w > 0 ? w - 1 : w + 1)), fr = smem_u32(flags + (w < NW - 1 ? w + 1 : w - 1));
        unsigned a, bb;
        do {
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(a) : "r"(fl) : "memory");
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(bb) : "r"(fr) : "memory");
        } while ((int)a < step + 1 || (int)bb < step + 1);
      } else if (XM == 4) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[b * NW + w]);
        const unsigned par = (step >> 1) & 1;
        mbar_test_wait(&bars[b * NW + (w > 0 ? w - 1 : w + 1)], par);
        mbar_test_wait(&bars[b * NW + (w < NW - 1 ? w + 1 : w - 1)], par);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[b * NW + w]);
        const unsigned par = (step >> 1) & 1;
        mbar_wait(&bars[b * NW + (w > 0 ? w - 1 : w + 1)], par);
        mbar_wait(&bars[b * NW + (w < NW - 1 ? w + 1 : w - 1)], par);
      }
      if (lane == 0 && (XM == 6 ? wc > 0 : w > 0)) {
        const double* src = edge + (b * NW + w - 1) * 12;
#pragma unroll
        for (int i = 0; i < 8; i += 2) { const double2 v = *reinterpret_cast<const double2*>(src + i); s[i] = v.x; s[i + 1] = v.y; }
      }
      if (lane == 31 && (XM == 6 ? wc < wpc - 1 : w < NW - 1)) {
        const double* src = edge + (b * NW + w + 1) * 12 + 8;
#pragma unroll
        for (int i = 0; i < 4; i += 2) { const double2 v = *reinterpret_cast<const double2*>(src + i); s[P - 4 + i] = v.x; s[P - 3 + i] = v.y; }
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (CM == 0) { xh[i] = s[i] + h2F; xd[i] = s[i] + dtF; }
      else { xh[i] = s[i] + c.h2F; xd[i] = s[i] + c.dtF; }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + t] = r;
}

// plain issue-rate probes: NOPS register operands per DFMA
template <int KIND>
__global__ void __launch_bounds__(384, 1) rate(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  double ra = a + threadIdx.x * 1e-12, rb = b + threadIdx.x * 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (KIND == 0) x[i] = fma(x[i], a, b);                         // 1 register operand (a, b uniform)
      else if (KIND == 1) x[i] = fma(x[i], ra, b);                   // 2 register operands, one shared
      else if (KIND == 2) x[i] = fma(x[i], ra, rb);                  // 3 register operands, two shared (reuse cache)
      else if (KIND == 3) x[i] = fma(x[(i + 5) & 15], x[(i + 9) & 15], x[i]);   // 3 distinct register operands
      else if (KIND == 4) x[i] = x[i] + x[(i + 5) & 15];             // DADD 2 distinct
      else if (KIND == 5) x[i] = fma(x[(i + 5) & 15], a, x[i]);      // 2 distinct + uniform
      else if (KIND == 6) x[i] = x[i] + b;                           // DADD 1 reg + uniform
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static int g_sms = 148;
static double* g_out;

template <int CM, int XM, int MAXT, int MINB, int NCH = 1, int P = 16>
void run(const char* name, int T, int ctas_per_sm, int stagger = 0) {
  const int steps = 400;
  const int NW = T / 32;
  const size_t smem = (2 * NW * 12 + 2 * NW) * sizeof(double) + 64;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = g_sms * ctas_per_sm * 2;
  const double dt = 0.01, F = 8.0;
  Coef c{0.5 * dt, dt / 3.0, dt / 6.0, dt, 0.5 * dt * F, dt * F};
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    probe<CM, XM, MAXT, MINB, NCH, P><<<grid, T, smem>>>(g_out, steps, dt, F, c, stagger);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double inst = (double)grid * T * P * steps * 17;
  const double rate = inst / (best * 1e-3) / (g_sms * 64.0) / 1.965e9;
  printf("%-34s P=%d CM=%d XM=%d NCH=%d stagger=%4d T=%3d x%d  %.3f ms  FP64 pipe %.1f %%  (%.2f TF by the 29-flop convention) %s\n", name, P, CM, XM, NCH, stagger, T,
         ctas_per_sm, best, 100 * rate, (double)grid * T * P * steps * 29 / (best * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
}

template <int KIND>
void run_rate(const char* name, int T) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    rate<KIND><<<g_sms, T>>>(g_out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double inst = (double)g_sms * T * 16 * iters;
  printf("rate %-44s T=%3d  %.3f ms  %.1f %% of 64 lanes/clk/SM @1965\n", name, T, best,
         100 * inst / (best * 1e-3) / (g_sms * 64.0) / 1.965e9);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  g_sms = p.multiProcessorCount;
  cudaMalloc(&g_out, sizeof(double) * 148 * 16 * 512);
  // warps per SMSP: how well do 1 / 2 / 3 warps fill the FP64 pipe?
  run<1, 1, 128, 1>("shuffles, 1 warp/SMSP", 128, 1);
  run<1, 1, 256, 1>("shuffles, 2 warps/SMSP", 256, 1);
  run<1, 1, 384, 1>("shuffles, 3 warps/SMSP", 384, 1);
  run<1, 0, 128, 1>("regs only, 1 warp/SMSP", 128, 1);
  run<1, 0, 256, 1>("regs only, 2 warps/SMSP", 256, 1);
  run<1, 2, 128, 1>("step halo sync, 1 warp/SMSP", 128, 1);
  run<1, 2, 256, 1>("step halo sync, 2 warps/SMSP", 256, 1);
  run<1, 6, 256, 1, 2>("step halo 2 chains x 4, 2 warps/SMSP", 256, 1);
  // more, smaller warps: P = 12 with 16 warps per SM
  run<1, 1, 512, 1, 1, 12>("shuffles P=12 4 warps/SMSP", 512, 1);
  run<1, 2, 512, 1, 1, 12>("step halo sync P=12 1 chain x 16", 512, 1);
  run<1, 6, 512, 1, 2, 12>("step halo P=12 2 chains x 8", 512, 1);
  run<1, 6, 512, 1, 4, 12>("step halo P=12 4 chains x 4", 512, 1);
  run<1, 6, 512, 1, 8, 12>("step halo P=12 8 chains x 2", 512, 1);
  run<1, 2, 128, 4, 1, 12>("step halo sync P=12 128x4", 128, 4);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
