// Standalone check + timing of the warp-dataflow forecast kernel (csrc/l96_forecast_v2.cu) against a plain
// CPU RK4 (textbook form), periodic and extended-layout inputs.  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I dataassimbench_b200/csrc -o build/l96_v2_test tools/l96_v2_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "experiments/l96_forecast_v2.cu"
#include "../dataassimbench_b200/csrc/l96_forecast_v3.cu"
namespace dab { bool pdl_enabled() { return false; } }
using namespace dab;

static void cpu_rk4(std::vector<double>& x, long long N, int steps, double dt, double F) {
  std::vector<double> k1(N), k2(N), k3(N), k4(N), y(N);
  auto rhs = [&](const std::vector<double>& v, std::vector<double>& d) {
    for (long long i = 0; i < N; ++i) {
      const double a = v[(i + 1) % N], b = v[(i - 2 + N) % N], c = v[(i - 1 + N) % N];
      d[i] = (a - b) * c - v[i] + F;
    }
  };
  for (int s = 0; s < steps; ++s) {
    rhs(x, k1);
    for (long long i = 0; i < N; ++i) y[i] = x[i] + 0.5 * dt * k1[i];
    rhs(y, k2);
    for (long long i = 0; i < N; ++i) y[i] = x[i] + 0.5 * dt * k2[i];
    rhs(y, k3);
    for (long long i = 0; i < N; ++i) y[i] = x[i] + dt * k3[i];
    rhs(y, k4);
    for (long long i = 0; i < N; ++i) x[i] += dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
  }
}

struct Bufs { double *in, *out, *s0, *s1; unsigned* flags; unsigned epoch = 0; };

// world ranks emulated one after the other: rank r owns columns [r n_loc, (r+1) n_loc) and gets the
// extended layout [8S | slab | 4S] cut from the ring
static int g_ver = 2;
static double check(long long N, int k, int S, int chunk, bool periodic, int world, int threads, bool verbose) {
  const double dt = 0.01, F = 8.0;
  std::vector<double> X((size_t)k * N);
  for (int m = 0; m < k; ++m)
    for (long long i = 0; i < N; ++i) X[(size_t)m * N + i] = 8.0 + 3.0 * sin(0.37 * i + m) + ((i * 7 + m * 13) % 11) * 0.1;
  std::vector<double> ref = X;
  for (int m = 0; m < k; ++m) {
    std::vector<double> row(ref.begin() + (size_t)m * N, ref.begin() + (size_t)(m + 1) * N);
    cpu_rk4(row, N, S, dt, F);
    std::copy(row.begin(), row.end(), ref.begin() + (size_t)m * N);
  }
  const long long n_loc = N / world;
  const long long ld_e = (8 * S + n_loc + 4 * S + 1) & ~1ll;
  const long long scr_ld = l96_v3_scratch_ld(n_loc, S);
  double *d_in, *d_out, *d_s0, *d_s1; unsigned* d_flags;
  const size_t in_elems = periodic ? (size_t)k * N : (size_t)k * ld_e;
  cudaMalloc(&d_in, in_elems * 8); cudaMalloc(&d_out, (size_t)k * n_loc * 8);
  cudaMalloc(&d_s0, (size_t)k * scr_ld * 8); cudaMalloc(&d_s1, (size_t)k * scr_ld * 8);
  const size_t nf = l96_v3_flag_count(n_loc, k, S);
  cudaMalloc(&d_flags, nf * 4); cudaMemset(d_flags, 0, nf * 4);
  unsigned long long* d_ctr; unsigned long long ctr_base = 0; cudaMalloc(&d_ctr, 8); cudaMemset(d_ctr, 0, 8);
  unsigned long long* d_ctr3; cudaMalloc(&d_ctr3, 16); cudaMemset(d_ctr3, 0, 16);
  double worst = 0.0, scale = 0.0;
  for (double v : ref) scale = fmax(scale, fabs(v));
  unsigned epoch = 0;
  for (int r = 0; r < world; ++r) {
    std::vector<double> hin(in_elems);
    if (periodic) hin = X;
    else
      for (int m = 0; m < k; ++m)
        for (long long e = 0; e < 8 * S + n_loc + 4 * S; ++e) {
          long long g = ((r * n_loc + e - 8 * S) % N + N) % N;
          hin[(size_t)m * ld_e + e] = X[(size_t)m * N + g];
        }
    cudaMemcpy(d_in, hin.data(), in_elems * 8, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xff, (size_t)k * n_loc * 8);
    cudaError_t e = g_ver == 2
        ? l96_v2_launch(d_in, d_out, d_s0, d_s1, periodic ? n_loc : scr_ld, d_flags, ++epoch, d_ctr, &ctr_base, n_loc, k, S, chunk, dt, F,
                        periodic ? N : ld_e, n_loc, periodic, periodic ? 0 : 8 * S, periodic ? 0 : -8ll * S,
                        periodic ? 0 : n_loc + 4ll * S, 148, 0)
        : l96_v3_launch(d_in, d_out, d_s0, d_s1, periodic ? n_loc : scr_ld, d_flags, ++epoch, d_ctr3, &ctr_base, n_loc, k, S, chunk, dt, F,
                        periodic ? N : ld_e, n_loc, periodic, periodic ? 0 : 8 * S, periodic ? 0 : -8ll * S,
                        periodic ? 0 : n_loc + 4ll * S, 148, 0);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1e9; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<double> hout((size_t)k * n_loc);
    cudaMemcpy(hout.data(), d_out, hout.size() * 8, cudaMemcpyDeviceToHost);
    for (int m = 0; m < k; ++m)
      for (long long i = 0; i < n_loc; ++i) {
        const double d = fabs(hout[(size_t)m * n_loc + i] - ref[(size_t)m * N + r * n_loc + i]);
        if (!(d <= worst)) worst = d;    // NaN-propagating max
      }
  }
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_s0); cudaFree(d_s1); cudaFree(d_flags);
  if (verbose) printf("v%d check N=%lld k=%d S=%d chunk=%d %s world=%d T=%d: max rel err %.3e\n", g_ver, N, k, S, chunk,
                      periodic ? "periodic" : "extended", world, threads, worst / scale);
  return worst / scale;
}

static void timeit(long long N, int k, int S, int chunk, int threads) {
  const double dt = 0.01, F = 8.0;
  const long long ld_e = (8 * S + N + 4 * S + 1) & ~1ll, scr_ld = l96_v3_scratch_ld(N, S);
  double *d_in, *d_out, *d_s0, *d_s1; unsigned* d_flags;
  cudaMalloc(&d_in, (size_t)k * ld_e * 8); cudaMalloc(&d_out, (size_t)k * N * 8);
  cudaMalloc(&d_s0, (size_t)k * scr_ld * 8); cudaMalloc(&d_s1, (size_t)k * scr_ld * 8);
  std::vector<double> h((size_t)k * ld_e);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 8.0 + 2.0 * sin(0.01 * i);
  cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t nf = l96_v3_flag_count(N, k, S);
  cudaMalloc(&d_flags, nf * 4); cudaMemset(d_flags, 0, nf * 4);
  unsigned long long* d_ctr; unsigned long long ctr_base = 0; cudaMalloc(&d_ctr, 8); cudaMemset(d_ctr, 0, 8);
  unsigned long long* d_ctr3; cudaMalloc(&d_ctr3, 16); cudaMemset(d_ctr3, 0, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned epoch = 0;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    if (g_ver == 2) l96_v2_launch(d_in, d_out, d_s0, d_s1, scr_ld, d_flags, ++epoch, d_ctr, &ctr_base, N, k, S, chunk, dt, F, ld_e, N, false, 8 * S, -8ll * S,
                  N + 4ll * S, 148, 0);
    else l96_v3_launch(d_in, d_out, d_s0, d_s1, scr_ld, d_flags, ++epoch, d_ctr3, &ctr_base, N, k, S, chunk, dt, F, ld_e, N, false, 8 * S, -8ll * S,
                  N + 4ll * S, 148, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("v%d time N=%lld k=%d S=%d chunk=%d T=%d: %.1f us  %.2f TFLOP/s (29-flop convention)  %s\n", g_ver, N, k, S, chunk, threads,
         best * 1e3, 29.0 * N * k * S / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_s0); cudaFree(d_s1); cudaFree(d_flags);
}

int main(int argc, char** argv) {
  for (g_ver = 3; g_ver >= 2; --g_ver) {
    const int c1 = g_ver == 3 ? 10 : 5, c2 = g_ver == 3 ? 7 : 4;
    double w = 0;
    w = fmax(w, check(1024, 4, 20, c1, true, 1, 384, true));
    w = fmax(w, check(1024, 16, 20, c1, false, 1, 384, true));
    w = fmax(w, check(600, 3, 7, c2, true, 1, 128, true));
    w = fmax(w, check(40, 20, 20, c1, true, 1, 384, true));
    w = fmax(w, check(36, 10, 20, c2, false, 1, 384, true));
    w = fmax(w, check(4096, 8, 20, c1, false, 4, 384, true));
    w = fmax(w, check(4096, 8, 32, c2, false, 2, 256, true));
    w = fmax(w, check(65536, 8, 20, c1, false, 1, 384, true));
    w = fmax(w, check(65536, 4, 20, c2, true, 1, 384, true));
    w = fmax(w, check(2000, 64, 3, c1, false, 1, 384, true));
    w = fmax(w, check(1 << 20, 2, 20, c1, false, 8, 384, true));
    if (g_ver == 3) { w = fmax(w, check(65536, 8, 20, 20, false, 1, 384, true)); w = fmax(w, check(8192, 64, 20, 5, false, 1, 384, true)); }
    printf("v%d worst %.3e %s\n", g_ver, w, w < 1e-11 ? "OK" : "FAIL");
  }
  if (argc > 1) return 0;
  g_ver = 3;
  for (int chunk : {20, 10, 7, 5, 4}) timeit(65536, 64, 20, chunk, 384);
  for (int chunk : {20, 10, 5}) timeit(8192, 64, 20, chunk, 384);      // C3 on 8 ranks
  for (int chunk : {20, 10, 5}) timeit(1 << 19, 64, 20, chunk, 384);   // C4 on 8 ranks
  for (int chunk : {20, 10, 5}) timeit(1 << 22, 64, 20, chunk, 384);   // C4
  for (int chunk : {10}) timeit(1 << 20, 128, 20, chunk, 384);  // C5
  return 0;
}
