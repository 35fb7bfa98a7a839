// Where does the forecast kernel's time go?  Same RK4 stage arithmetic as csrc/l96_forecast.cu
// (17 FP64 instructions per point and step, P = 8 points per thread), with the neighbour exchange
// progressively switched on (MODE 3: warp shuffles instead):  MODE 0: neighbours from the thread's own registers (no shared memory,
// no barrier) -> the instruction mix's own FP64-pipe ceiling;  MODE 1: + one __syncthreads per
// stage;  MODE 2: + shared-memory edge exchange (the real kernel's inner loop).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) probe(double* out, int n_steps, double dt, double F) {
  constexpr int P = 8;
  extern __shared__ double smem[];
  const int T = blockDim.x, t = threadIdx.x;
  double2* e_last2 = reinterpret_cast<double2*>(smem);
  double* e_first = smem + 4 * T;
  double s[P], xh[P], xd[P], acc[P];
  const double h2 = 0.5 * dt, h6 = dt / 6.0, h3 = dt / 3.0, h2F = h2 * F, dtF = dt * F;
#pragma unroll
  for (int i = 0; i < P; ++i) { s[i] = 8.0 + 0.01 * ((t * P + i) % 37) + blockIdx.x * 1e-6; xh[i] = s[i] + h2F; xd[i] = s[i] + dtF; }
  const int tl = t > 0 ? t - 1 : 0, tr = t < T - 1 ? t + 1 : T - 1;
  int buf = 0;
  if (MODE == 2) { e_first[t] = s[0]; e_last2[t] = make_double2(s[P - 2], s[P - 1]); }
  for (int step = 0; step < n_steps; ++step) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      double2 L; double R;
      if (MODE >= 1) __syncthreads();
      if (MODE == 2) { L = e_last2[buf * T + tl]; R = e_first[buf * T + tr]; buf ^= 1; }
      else if (MODE == 3) {     // warp shuffles: no shared memory, no barrier
        L.x = __shfl_up_sync(0xffffffffu, s[P - 2], 1); L.y = __shfl_up_sync(0xffffffffu, s[P - 1], 1);
        R = __shfl_down_sync(0xffffffffu, s[0], 1);
      }
      else { L = make_double2(s[P - 2], s[P - 1]); R = s[0]; }
      const double ca = (st == 2) ? dt : h2, cw = (st == 0 || st == 3) ? h6 : h3;
      double kk[P];
      kk[0] = fma(s[1] - L.x, L.y, -s[0]);
      kk[1] = fma(s[2] - L.y, s[0], -s[1]);
      kk[P - 1] = fma(R - s[P - 3], s[P - 2], -s[P - 1]);
      kk[P - 2] = fma(s[P - 1] - s[P - 4], s[P - 3], -s[P - 2]);
      if (MODE == 2) {
        double n0, n2, n1;
        if (st < 3) { n0 = fma(ca, kk[0], st == 2 ? xd[0] : xh[0]); n2 = fma(ca, kk[P - 2], st == 2 ? xd[P - 2] : xh[P - 2]); n1 = fma(ca, kk[P - 1], st == 2 ? xd[P - 1] : xh[P - 1]); }
        else { n0 = fma(cw, kk[0], acc[0]); n2 = fma(cw, kk[P - 2], acc[P - 2]); n1 = fma(cw, kk[P - 1], acc[P - 1]); }
        e_first[buf * T + t] = n0; e_last2[buf * T + t] = make_double2(n2, n1);
      }
#pragma unroll
      for (int i = 2; i < P - 2; ++i) kk[i] = fma(s[i + 1] - s[i - 2], s[i - 1], -s[i]);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (st == 0) acc[i] = fma(cw, kk[i], xd[i]);
        else if (st < 3) acc[i] = fma(cw, kk[i], acc[i]);
        if (st < 3) s[i] = fma(ca, kk[i], st == 2 ? xd[i] : xh[i]);
        else { s[i] = fma(cw, kk[i], acc[i]); xh[i] = s[i] + h2F; xd[i] = s[i] + dtF; }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) r += s[i];
  out[blockIdx.x * T + t] = r;
}

template <int MODE, int MAXT, int MINB>
void run(const char* name, int T, int ctas_per_sm, int sms, double* out) {
  const int steps = 400;
  const size_t smem = 6 * T * sizeof(double);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * ctas_per_sm * 4;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    probe<MODE, MAXT, MINB><<<grid, T, smem>>>(out, steps, 0.01, 8.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double inst = (double)grid * T * 8 * steps * 17;           // FP64 lane-instructions
  const double rate = inst / (ms * 1e-3) / (sms * 64.0) / 1.965e9;  // fraction of 64 lanes/clk/SM at 1965 MHz
  printf("%-40s T=%3d ctas/SM=%d  %.3f ms  FP64 pipe %.1f %%  (%.2f TFLOP/s by the 29-flop convention)\n", name, T, ctas_per_sm, ms,
         100 * rate, (double)grid * T * 8 * steps * 29 / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 4 * 512);
  run<0, 256, 3>("mode0 registers only (80 regs)", 256, 3, sms, out);
  run<1, 256, 3>("mode1 + barrier per stage", 256, 3, sms, out);
  run<2, 256, 3>("mode2 + smem edge exchange", 256, 3, sms, out);
  run<3, 256, 3>("mode3 warp shuffles 256x3", 256, 3, sms, out);
  run<3, 256, 2>("mode3 warp shuffles 256x2 (<=128 regs)", 256, 2, sms, out);
  run<3, 512, 1>("mode3 warp shuffles 512x1", 512, 1, sms, out);
  run<0, 256, 1>("mode0 256x1 (<=255 regs)", 256, 1, sms, out);
  run<0, 256, 2>("mode0 (<=128 regs)", 256, 2, sms, out);
  run<2, 256, 2>("mode2 (<=128 regs)", 256, 2, sms, out);
  run<0, 512, 1>("mode0 512 threads", 512, 1, sms, out);
  run<2, 512, 1>("mode2 512 threads", 512, 1, sms, out);
  run<2, 384, 2>("mode2 384x2", 384, 2, sms, out);
  run<2, 256, 3>("mode2 128x6", 128, 6, sms, out);
  run<0, 256, 3>("mode0 128x6", 128, 6, sms, out);
  run<0, 256, 3>("mode0 256x1", 256, 1, sms, out);
  run<0, 256, 3>("mode0 128x1", 128, 1, sms, out);
  run<0, 256, 3>("mode0 32x1 (one warp per SMSP? no: one warp per SM)", 32, 1, sms, out);
  run<0, 256, 3>("mode0 128x1", 128, 1, sms, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
