"""Oracle: Lorenz-96 tendencies and integrators (TEST INFRASTRUCTURE ONLY).

Follows the reference:
  * rhs                -- dabench/data/_lorenz96.py:119-151
  * default x0 tables  -- dabench/data/_lorenz96.py:66-117
  * time grid          -- dabench/data/_utils.py:44  (t = arange(0, t_final - dt/2, dt))
  * generate()         -- dabench/data/_data.py:121-125,161-176

Two integrators are provided:
  * `generate_dopri5`  -- the reference-faithful adaptive Dormand-Prince 5(4)
                          (`jax.experimental.ode.odeint`, restated in odeint_dopri5.py)
  * `rk4_forecast`     -- classical fixed-step RK4: the specification of the CUDA
                          forecast kernel (north star: "Lorenz96 RK4 ensemble forecast").
"""
import numpy as np

from .odeint_dopri5 import odeint as _odeint_dopri5

# dabench/data/_lorenz96.py:66-96 (the tables are float32 literals in the reference;
# with jax x64 enabled jnp.array() of python floats gives exactly these doubles).
X0_DEFAULTS = {
    0.05: {
        5: [3.552747, 5.173696, 5.4161925, 5.1463985, -0.62878036],
        6: [-1.277197, 0.5191239, 2.1629903, 4.641832, 5.013882, -3.1474972],
        36: [0.90061724, 2.2108543, 3.3563306, 7.0470520, 7.3828993, -2.2906365,
             1.6358340, 4.5246205, -0.8536633, 2.2018400, 2.5094680, 5.6148005,
             -1.7163916, -3.5827417, 0.22293478, 1.8138107, 3.7354333, 5.9006715,
             -4.6722836, 0.4664867, 0.36800075, 7.7004447, 3.0569422, -1.7238870,
             -2.1296368, 1.6388168, 5.1955190, 4.7863874, 0.8382774, -4.0938597,
             0.5181451, 1.2503184, 6.0076460, 7.1161866, -3.2190716, -2.3532054],
    },
    0.01: {
        5: [1.7660924, 0.29829425, 0.3133399, 4.521974, 8.680367],
        6: [4.321441, 9.965719, 3.7177398, 1.1666082, -0.26569614, 0.3100199],
        36: [1.0433275, 4.401105, 6.32247, 4.9386244, 3.4934044, 2.3050377,
             -1.1888806, 0.83992016, 2.52427, 10.146448, 1.5634892, 4.289391,
             5.6141367, 5.9368305, -1.291068, -2.2805283, 3.365577, 6.793376,
             0.41429618, 3.4392211, 5.739969, -3.6751282, 1.2353066, 0.04940181,
             1.9196223, 6.1237254, 3.356311, -2.9513268, 1.8416065, 3.6102607,
             9.518854, 2.4801574, -0.44426048, 5.1087484, -1.305786, -0.79453385],
    },
}


def default_x0(system_dim, delta_t):
    """dabench/data/_lorenz96.py:98-117."""
    table = X0_DEFAULTS.get(delta_t, X0_DEFAULTS[0.01])
    if system_dim in table:
        return np.array(table[system_dim], dtype=np.float64)
    x0 = np.zeros(system_dim)
    x0[0] = 0.01
    return x0


def rhs(x, F=8.0):
    """dx_i = (x_{i+1} - x_{i-2}) * x_{i-1} - x_i + F, periodic.

    dabench/data/_lorenz96.py:138-149; evaluation order ((d*xm1) - x) + F as there.
    Works on the last axis, so an ensemble [k, N] is advanced member-wise.
    """
    x = np.asarray(x, dtype=np.float64)
    d = np.roll(x, -1, axis=-1) - np.roll(x, 2, axis=-1)
    dx = d * np.roll(x, 1, axis=-1) - x
    dx = dx + F
    return dx


def rk4_step(x, dt, F=8.0):
    """One classical RK4 step (kernel specification)."""
    k1 = rhs(x, F)
    k2 = rhs(x + (0.5 * dt) * k1, F)
    k3 = rhs(x + (0.5 * dt) * k2, F)
    k4 = rhs(x + dt * k3, F)
    return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def rk4_forecast(x0, n_steps, dt, F=8.0, return_traj=False):
    """Advance `n_steps` RK4 steps.  With return_traj the result has n_steps+1
    points including x0, matching `generate(n_steps+1)` of the reference."""
    x = np.array(x0, dtype=np.float64, copy=True)
    if return_traj:
        traj = np.empty((n_steps + 1,) + x.shape)
        traj[0] = x
    for s in range(n_steps):
        x = rk4_step(x, dt, F)
        if return_traj:
            traj[s + 1] = x
    return (x, traj) if return_traj else x


def time_grid(n_steps, delta_t):
    """dabench/data/_data.py:122-123 + dabench/data/_utils.py:44."""
    t_final = n_steps * delta_t
    return np.arange(0.0, t_final - delta_t / 2, delta_t)


def generate_dopri5(x0, n_steps, delta_t, F=8.0, rtol=1.4e-8, atol=1.4e-8):
    """Reference-faithful `Lorenz96.generate(n_steps=..., x0=...)` for one member:
    returns [n_steps, N] including the initial state (dabench/data/_data.py:161-176)."""
    t = time_grid(n_steps, delta_t)
    return _odeint_dopri5(lambda y, _t: rhs(y, F), np.asarray(x0, dtype=np.float64), t,
                          rtol=rtol, atol=atol)


def generate_rk4(x0, n_steps, delta_t, F=8.0):
    """Same output contract as generate_dopri5 but fixed-step RK4."""
    _, traj = rk4_forecast(x0, n_steps - 1, delta_t, F, return_traj=True)
    return traj
