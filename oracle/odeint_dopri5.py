"""Oracle: adaptive Dormand-Prince 5(4) integrator (TEST INFRASTRUCTURE ONLY).

The reference integrates every model with `jax.experimental.ode.odeint`
(dabench/data/_utils.py:8,51).  JAX is a third-party dependency that is absent
from /root/reference and from this image (`pyproject.toml:21` leaves it
unpinned), so its published algorithm is restated here in NumPy:

  * Dormand-Prince 5(4) tableau with FSAL,
  * initial step from Hairer/Norsett/Wanner Sec. II.4 (order 4),
  * per attempted step: embedded error, ratio = sqrt(mean((err/(atol+rtol*max(|y0|,|y1|)))^2)),
    accept iff ratio <= 1, next dt = dt*min(10, max(0.9*ratio^(-1/5), dfactor)),
    dfactor = 1 if ratio < 1 else 0.2 (dt*10 if ratio == 0),
  * steps are NOT clipped to output times; outputs come from the 4th-order
    interpolating polynomial fitted to (y0, y1, y_mid, f0, f1) of the last
    accepted step,
  * defaults rtol = atol = 1.4e-8, mxstep = inf, hmax = inf (the reference passes none).

Anchor: with this integrator the oracle reproduces the reference's end-to-end
ETKF golden vectors (tests/dacycler_etkf_test.py:89-102) and agrees with
SciPy's independent RK45 to ~1e-8 (tests/test_oracle_golden.py).
"""
import numpy as np

_ALPHA = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0, 0.0])
_BETA = np.array([
    [1 / 5, 0, 0, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0, 0],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
])
_C_SOL = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])
_C_ERR = np.array([
    35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0])
_C_MID = np.array([
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])


def _initial_step_size(fun, t0, y0, order, rtol, atol, f0):
    scale = atol + np.abs(y0) * rtol
    d0 = np.linalg.norm(y0 / scale)
    d1 = np.linalg.norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = fun(y1, t0 + h0)
    d2 = np.linalg.norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1.0))
    return min(100.0 * h0, h1)


def _rk_step(fun, y0, f0, t0, dt):
    k = np.zeros((7,) + y0.shape)
    k[0] = f0
    for i in range(1, 7):
        ti = t0 + dt * _ALPHA[i - 1]
        yi = y0 + dt * np.tensordot(_BETA[i - 1], k, axes=1)
        k[i] = fun(yi, ti)
    y1 = dt * np.tensordot(_C_SOL, k, axes=1) + y0
    y1_err = dt * np.tensordot(_C_ERR, k, axes=1)
    return y1, k[-1], y1_err, k


def _interp_fit(y0, y1, k, dt):
    y_mid = y0 + dt * np.tensordot(_C_MID, k, axes=1)
    dy0, dy1 = k[0], k[-1]
    a = -2.0 * dt * dy0 + 2.0 * dt * dy1 - 8.0 * y0 - 8.0 * y1 + 16.0 * y_mid
    b = 5.0 * dt * dy0 - 3.0 * dt * dy1 + 18.0 * y0 + 14.0 * y1 - 32.0 * y_mid
    c = -4.0 * dt * dy0 + dt * dy1 - 11.0 * y0 - 5.0 * y1 + 16.0 * y_mid
    d = dt * dy0
    e = y0
    return (a, b, c, d, e)


def _optimal_step(last_step, ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5.0):
    if ratio == 0:
        return last_step * ifactor
    dfac = 1.0 if ratio < 1 else dfactor
    factor = min(ifactor, max(ratio ** (-1.0 / order) * safety, dfac))
    return last_step * factor


def odeint(fun, y0, ts, rtol=1.4e-8, atol=1.4e-8, mxstep=np.inf, hmax=np.inf):
    """y(ts) for dy/dt = fun(y, t); returns [len(ts), *y0.shape] with row 0 = y0."""
    y0 = np.asarray(y0, dtype=np.float64)
    ts = np.asarray(ts, dtype=np.float64)
    out = np.empty((len(ts),) + y0.shape)
    out[0] = y0
    f = fun(y0, ts[0])
    dt = min(max(_initial_step_size(fun, ts[0], y0, 4, rtol, atol, f), 0.0), hmax)
    y, t, last_t = y0, ts[0], ts[0]
    coeff = (y0, y0, y0, y0, y0)
    for j in range(1, len(ts)):
        target = ts[j]
        i = 0
        while t < target and i < mxstep and dt > 0:
            ny, nf, nerr, k = _rk_step(fun, y, f, t, dt)
            tol = atol + rtol * np.maximum(np.abs(y), np.abs(ny))
            ratio = float(np.sqrt(np.mean((nerr / tol) ** 2)))
            new_coeff = _interp_fit(y, ny, k, dt)
            new_dt = min(max(_optimal_step(dt, ratio), 0.0), hmax)
            if ratio <= 1.0:
                last_t, t = t, t + dt
                y, f, coeff = ny, nf, new_coeff
            dt = new_dt
            i += 1
        rel = (target - last_t) / (t - last_t)
        a, b, c, d, e = coeff
        out[j] = (((a * rel + b) * rel + c) * rel + d) * rel + e
    return out
