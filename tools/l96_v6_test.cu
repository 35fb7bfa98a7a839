// Standalone check + timing of the forecast kernel with chains handed out inside the SM (csrc/l96_forecast_v6.cu) against a plain CPU RK4
// (textbook form), periodic and extended-layout inputs, and against the first design (csrc/l96_forecast.cu) bit for
// bit.  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I dataassimbench_b200/csrc -o build/l96_v6_test tools/l96_v6_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../dataassimbench_b200/csrc/l96_forecast_v6.cu"
#include "../dataassimbench_b200/csrc/l96_forecast.cu"
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

static int g_slots = 444, g_variant = 0;

// world ranks emulated one after the other: rank r owns columns [r n_loc, (r+1) n_loc) and gets the
// extended layout [8S | slab | 4S] cut from the ring
static double check(long long N, int k, int S, bool periodic, int world, int slots, bool verbose) {
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
  double *d_in, *d_out, *d_out1;
  const size_t in_elems = periodic ? (size_t)k * N : (size_t)k * ld_e;
  cudaMalloc(&d_in, in_elems * 8); cudaMalloc(&d_out, (size_t)k * n_loc * 8); cudaMalloc(&d_out1, (size_t)k * n_loc * 8);
  L96V6Plan plan;
  if (!l96_v6_plan(n_loc, k, S, slots, g_variant, &plan)) { printf("plan failed\n"); return 1e9; }
  int *d_jobs, *d_chains; cudaMalloc(&d_jobs, plan.tiles.size() * 4); cudaMalloc(&d_chains, plan.chains.size() * 4);
  cudaMemcpy(d_jobs, plan.tiles.data(), plan.tiles.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_chains, plan.chains.data(), plan.chains.size() * 4, cudaMemcpyHostToDevice);
  double worst = 0.0, scale = 0.0;
  long long bit_diff = 0;
  for (double v : ref) scale = fmax(scale, fabs(v));
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
    cudaMemset(d_out1, 0xff, (size_t)k * n_loc * 8);
    cudaError_t e = l96_v6_launch(d_in, d_out, d_jobs, d_chains, plan.grid, plan.tile_stride, plan.chain_stride, plan.hist_t, g_variant, n_loc, S, dt, F,
                                  periodic ? N : ld_e, n_loc, periodic, periodic ? 0 : 8 * S, periodic ? 0 : -8ll * S,
                                  periodic ? 0 : n_loc + 4ll * S, 0);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1e9; }
    e = l96_forecast_launch(d_in, d_out1, nullptr, n_loc, k, S, dt, F, periodic ? N : ld_e, n_loc, periodic,
                            periodic ? 0 : 8 * S, periodic ? 0 : -8ll * S, periodic ? 0 : n_loc + 4ll * S, 0, 148, 0, 0);
    if (e != cudaSuccess) { printf("launch v1: %s\n", cudaGetErrorString(e)); return 1e9; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<double> hout((size_t)k * n_loc), hout1((size_t)k * n_loc);
    cudaMemcpy(hout.data(), d_out, hout.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hout1.data(), d_out1, hout1.size() * 8, cudaMemcpyDeviceToHost);
    for (int m = 0; m < k; ++m)
      for (long long i = 0; i < n_loc; ++i) {
        const double d = fabs(hout[(size_t)m * n_loc + i] - ref[(size_t)m * N + r * n_loc + i]);
        if (!(d <= worst)) worst = d;    // NaN-propagating max
      }
    bit_diff += memcmp(hout.data(), hout1.data(), hout.size() * 8) != 0;
  }
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_out1); cudaFree(d_jobs); cudaFree(d_chains);
  if (verbose) printf("v6 check N=%lld k=%d S=%d %s world=%d slots=%d: grid %d x <= %d tiles, max rel err %.3e, %s first design\n", N, k, S,
                      periodic ? "periodic" : "extended", world, slots, plan.grid, plan.tile_stride, worst / scale,
                      bit_diff ? "DIFFERS from" : "bit-identical to");
  return bit_diff ? 1.0 : worst / scale;
}

static void timeit(long long N, int k, int S, int slots) {
  const double dt = 0.01, F = 8.0;
  const long long ld_e = (8 * S + N + 4 * S + 1) & ~1ll;
  double *d_in, *d_out;
  cudaMalloc(&d_in, (size_t)k * ld_e * 8); cudaMalloc(&d_out, (size_t)k * N * 8);
  std::vector<double> h((size_t)k * ld_e);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 8.0 + 2.0 * sin(0.01 * i);
  cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  L96V6Plan plan;
  if (!l96_v6_plan(N, k, S, slots, g_variant, &plan)) { printf("plan failed\n"); return; }
  int *d_jobs, *d_chains; cudaMalloc(&d_jobs, plan.tiles.size() * 4); cudaMalloc(&d_chains, plan.chains.size() * 4);
  cudaMemcpy(d_jobs, plan.tiles.data(), plan.tiles.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_chains, plan.chains.data(), plan.chains.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    l96_v6_launch(d_in, d_out, d_jobs, d_chains, plan.grid, plan.tile_stride, plan.chain_stride, plan.hist_t, g_variant, N, S, dt, F, ld_e, N, false, 8 * S,
                  -8ll * S, N + 4ll * S, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("v6 time N=%lld k=%d S=%d slots=%d (grid %d x <= %d tiles): %.1f us  %.2f TFLOP/s (29-flop convention)  %s\n", N, k, S, slots,
         plan.grid, plan.tile_stride, best * 1e3, 29.0 * N * k * S / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_jobs); cudaFree(d_chains);
}

int main(int argc, char** argv) {
  g_slots = 148;
  if (getenv("V6_VARIANT")) g_variant = atoi(getenv("V6_VARIANT"));
  printf("variant %d\n", g_variant);
  
  double w = 0;
  if (getenv("DAB_SKIP_CHECK")) goto timing;
  w = fmax(w, check(1024, 4, 20, true, 1, g_slots, true));
  w = fmax(w, check(1024, 16, 20, false, 1, g_slots, true));
  w = fmax(w, check(600, 3, 7, true, 1, g_slots, true));
  w = fmax(w, check(40, 20, 20, true, 1, g_slots, true));
  w = fmax(w, check(36, 10, 20, false, 1, g_slots, true));
  w = fmax(w, check(4096, 8, 20, false, 4, g_slots, true));
  w = fmax(w, check(4096, 8, 32, false, 2, g_slots, true));
  w = fmax(w, check(65536, 8, 20, false, 1, g_slots, true));
  w = fmax(w, check(65536, 8, 20, false, 1, 7, true));        // long chains that run from one member into the next
  w = fmax(w, check(65536, 4, 20, true, 1, 5, true));
  w = fmax(w, check(65536, 4, 19, true, 1, 5, true));         // 4 S not a multiple of 16
  w = fmax(w, check(65536, 4, 1, false, 1, 3, true));
  w = fmax(w, check(2000, 64, 3, false, 1, g_slots, true));
  w = fmax(w, check(1 << 20, 2, 20, false, 8, g_slots, true));
  w = fmax(w, check(1 << 20, 2, 20, true, 1, 16, true));
  printf("v6 worst %.3e %s\n", w, w < 1e-11 ? "OK" : "FAIL");
  if (argc > 1) return 0;
timing:
  timeit(65536, 64, 20, g_slots);
  timeit(65536, 64, 20, 147);
  timeit(32768, 64, 20, g_slots);      // C3 on 2 ranks
  timeit(8192, 64, 20, g_slots);       // C3 on 8 ranks
  timeit(1 << 19, 64, 20, g_slots);    // C4 on 8 ranks
  timeit(1 << 22, 64, 20, g_slots);    // C4
  timeit(1 << 20, 128, 20, g_slots);   // C5
  return 0;
}
