// per-phase clock64 sums of one warp of the warp-dataflow forecast kernel (debug build)
#define V2_DEBUG_TIMING
#include <cstdio>
#include <vector>
#include <cmath>
#include "experiments/l96_forecast_v2.cu"
namespace dab { bool pdl_enabled() { return false; } }
using namespace dab;
int main() {
  const long long N = 65536; const int k = 64, S = 20, chunk = 5; const double dt = 0.01, F = 8.0;
  const long long ld_e = (8 * S + N + 4 * S + 1) & ~1ll, scr_ld = l96_v2x_scratch_ld(N, S);
  double *d_in, *d_out, *d_s0, *d_s1; unsigned* d_flags;
  cudaMalloc(&d_in, (size_t)k * ld_e * 8); cudaMalloc(&d_out, (size_t)k * N * 8);
  cudaMalloc(&d_s0, (size_t)k * scr_ld * 8); cudaMalloc(&d_s1, (size_t)k * scr_ld * 8);
  std::vector<double> h((size_t)k * ld_e);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 8.0 + 2.0 * sin(0.01 * i);
  cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t nf = l96_v2x_flag_count(N, k, S);
  cudaMalloc(&d_flags, nf * 4); cudaMemset(d_flags, 0, nf * 4);
  unsigned long long* d_ctr; unsigned long long ctr_base = 0; cudaMalloc(&d_ctr, 8); cudaMemset(d_ctr, 0, 8);
  unsigned epoch = 0;
  for (int rep = 0; rep < 3; ++rep) {
    long long z[8] = {0};
    cudaMemcpyToSymbol(g_v2_t, z, sizeof(z));
    l96_v2_launch(d_in, d_out, d_s0, d_s1, scr_ld, d_flags, ++epoch, d_ctr, &ctr_base, N, k, S, chunk, dt, F, ld_e, N, false, 8 * S, -8ll * S,
                  N + 4ll * S, 148, 0);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(z, g_v2_t, sizeof(z));
    long long cn[4]; cudaMemcpyFromSymbol(cn, g_v2_cnt, sizeof(cn)); printf("jobs %lld not-fetched %lld many %lld last %lld\n", cn[0], cn[1], cn[2], cn[3]);
    long long zz[4] = {0}; cudaMemcpyToSymbol(g_v2_cnt, zz, sizeof(zz));
    long long tot = 0; for (int i = 0; i < 8; ++i) tot += z[i];
    printf("cycles: late-fetch %lld | wait+load %lld | claim+decode %lld | compute %lld | stage+store %lld | release %lld | loop %lld | total %lld (%.1f us)\n",
           z[5], z[0], z[1], z[2], z[3], z[4], z[7], tot, tot / 1965.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
