"""Oracle: Lorenz-63 tendencies and integrators (TEST INFRASTRUCTURE ONLY).

Follows the reference:
  * rhs, parameters, default x0 -- dabench/data/_lorenz63.py:37-90
  * time grid / generate()      -- dabench/data/_utils.py:44, dabench/data/_data.py:121-125,161-176

`generate_dopri5` is the reference-faithful adaptive integrator (jax.experimental.ode.odeint,
restated in odeint_dopri5.py); `rk4_forecast` is the specification of the CUDA kernel
(`dab_l63_forecast_f64`).  Pinned by the reference's known answer
tests/model_base_test.py:18-30 (tests/test_oracle_golden.py).
"""
import numpy as np

from .odeint_dopri5 import odeint as _odeint_dopri5

X0_DEFAULT = np.array([-10.0, -15.0, 21.3])        # dabench/data/_lorenz63.py:42


def rhs(x, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
    """dabench/data/_lorenz63.py:84-88; works on the last axis (an ensemble [k, 3] member-wise)."""
    x = np.asarray(x, dtype=np.float64)
    x0, x1, x2 = x[..., 0], x[..., 1], x[..., 2]
    return np.stack([sigma * (x1 - x0), rho * x0 - x1 - x0 * x2, x0 * x1 - beta * x2], axis=-1)


def rk4_step(x, dt, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
    k1 = rhs(x, sigma, rho, beta)
    k2 = rhs(x + (0.5 * dt) * k1, sigma, rho, beta)
    k3 = rhs(x + (0.5 * dt) * k2, sigma, rho, beta)
    k4 = rhs(x + dt * k3, sigma, rho, beta)
    return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def rk4_forecast(x0, n_steps, dt, sigma=10.0, rho=28.0, beta=8.0 / 3.0, return_traj=False):
    """Advance n_steps RK4 steps; with return_traj the trajectory has n_steps+1 points including x0."""
    x = np.array(x0, dtype=np.float64, copy=True)
    traj = np.empty((n_steps + 1,) + x.shape) if return_traj else None
    if return_traj:
        traj[0] = x
    for s in range(n_steps):
        x = rk4_step(x, dt, sigma, rho, beta)
        if return_traj:
            traj[s + 1] = x
    return (x, traj) if return_traj else x


def generate_rk4(x0, n_points, dt, **par):
    """`generate(n_steps=n_points)` with the RK4 integrator: n_points states including x0."""
    return rk4_forecast(x0, n_points - 1, dt, return_traj=True, **par)[1]


def generate_dopri5(x0, n_points, dt, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
    """`Lorenz63.generate(n_steps=n_points)` as the reference integrates it (adaptive Dormand-Prince)."""
    t = np.arange(0.0, n_points * dt - dt / 2, dt)
    return _odeint_dopri5(lambda y, tt: rhs(y, sigma, rho, beta), np.asarray(x0, dtype=np.float64), t)
