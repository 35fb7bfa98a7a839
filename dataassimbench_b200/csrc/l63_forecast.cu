// Lorenz-63 RK4 ensemble forecast (another generator behind the forecast ABI).
//
// Replaces, for `model_obj = Lorenz63`, the same chain K1 replaces for Lorenz-96:
//   _step_forecast -> Model.forecast -> Data.generate -> integrate/odeint -> Lorenz63.rhs
//   (dabench/dacycler/_etkf.py:64-80, dabench/data/_data.py:86-211, dabench/data/_lorenz63.py:70-90)
// with classical fixed-step RK4 at the model's delta_t (the integrator of the forecast kernels).
// A member is three doubles, so there is nothing to tile: one thread owns one member for the whole
// window, its state never leaves registers, and the only memory traffic is 24 B in, 24 B out (plus
// 24 B per step when the trajectory is requested).  72 FP64 instructions per member-step.
#include "common.cuh"
#include "kernels.h"

namespace dab {

struct L63Tend { double x, y, z; };

__device__ __forceinline__ L63Tend l63_rhs(double x, double y, double z, double sigma, double rho, double beta) {
  L63Tend t;
  t.x = sigma * (y - x);
  t.y = fma(rho, x, -y) - x * z;
  t.z = fma(x, y, -beta * z);
  return t;
}

__global__ void __launch_bounds__(128)
l63_rk4_kernel(const double* __restrict__ x_in, double* __restrict__ x_out, double* __restrict__ traj,
               long long k, int n_steps, double dt, double sigma, double rho, double beta) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= k) return;
  double x = x_in[3 * j], y = x_in[3 * j + 1], z = x_in[3 * j + 2];
  const double h2 = 0.5 * dt, h6 = dt / 6.0;
  for (int s = 0; s < n_steps; ++s) {
    const L63Tend k1 = l63_rhs(x, y, z, sigma, rho, beta);
    const L63Tend k2 = l63_rhs(fma(h2, k1.x, x), fma(h2, k1.y, y), fma(h2, k1.z, z), sigma, rho, beta);
    const L63Tend k3 = l63_rhs(fma(h2, k2.x, x), fma(h2, k2.y, y), fma(h2, k2.z, z), sigma, rho, beta);
    const L63Tend k4 = l63_rhs(fma(dt, k3.x, x), fma(dt, k3.y, y), fma(dt, k3.z, z), sigma, rho, beta);
    x = fma(h6, (k1.x + 2.0 * k2.x) + (2.0 * k3.x + k4.x), x);
    y = fma(h6, (k1.y + 2.0 * k2.y) + (2.0 * k3.y + k4.y), y);
    z = fma(h6, (k1.z + 2.0 * k2.z) + (2.0 * k3.z + k4.z), z);
    if (traj != nullptr) {
      double* t = traj + ((long long)s * k + j) * 3;
      t[0] = x; t[1] = y; t[2] = z;
    }
  }
  x_out[3 * j] = x; x_out[3 * j + 1] = y; x_out[3 * j + 2] = z;
}

cudaError_t l63_forecast_launch(const double* x_in, double* x_out, double* traj, long long k, int n_steps,
                                double dt, double sigma, double rho, double beta, cudaStream_t st) {
  if (k <= 0) return cudaSuccess;
  l63_rk4_kernel<<<(unsigned)((k + 127) / 128), 128, 0, st>>>(x_in, x_out, traj, k, n_steps, dt, sigma, rho, beta);
  return cudaGetLastError();
}

}  // namespace dab
