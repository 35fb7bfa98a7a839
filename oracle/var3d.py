"""Oracle: the 3D-Var cycler (TEST INFRASTRUCTURE ONLY).

Restates `dabench.dacycler.Var3D._cycle_obsop` (dabench/dacycler/_var3d.py:52-112) literally --
dense H, dense R, B = I by default, the "preconditioned with B" system
    (I + B H^T R^-1 H) xa = xb + B H^T R^-1 y
solved with conjugate gradients from x0 = xb (tol 1e-5, maxiter 1000) -- and the part of
`DACycler.cycle` it runs under (dabench/dacycler/_dacycler.py:110-141, 186-290: same window
bookkeeping as the ETKF, no ensemble dimension).  `cg` follows jax.scipy.sparse.linalg.cg
(unpreconditioned: stop when ||r||^2 <= max(tol^2 ||b||^2, atol^2) or after maxiter steps).

`analysis_diagonal` is the closed form for the default operators (index-gather H, R = sd^2 I,
B = I): the system is diagonal, xa_i = (xb_i + s_i / sd^2) / (1 + n_i / sd^2) with n_i, s_i the
number and the sum of the valid observations of grid point i.  SURVEY.md 8f rank 4; pinned by
/root/reference/tests/dacycler_var3d_test.py:56-84 (tests/test_oracle_golden.py).
"""
import numpy as np

from . import windows as _win
from .etkf import calc_default_H, calc_default_R


def cg(A, b, x0, tol=1e-5, atol=0.0, maxiter=1000):
    """jax.scipy.sparse.linalg.cg with M = I, dense A."""
    b = np.asarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    atol2 = max(tol ** 2 * float(b @ b), atol ** 2)
    r = b - A @ x
    p = r.copy()
    gamma = float(r @ r)
    k = 0
    while gamma > atol2 and k < maxiter:
        Ap = A @ p
        alpha = gamma / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        gamma_new = float(r @ r)
        p = r + (gamma_new / gamma) * p
        gamma = gamma_new
        k += 1
    return x, k


def analysis_literal(xb, y, H, R, B=None):
    """dabench/dacycler/_var3d.py:84-110 (masks already applied to H by the caller)."""
    xb = np.asarray(xb, dtype=np.float64)
    n = xb.size
    if B is None:
        B = np.identity(n)                                     # _dacycler.py:74-76
    Rinv = np.linalg.inv(R)
    BHtRinv = (B @ H.T) @ Rinv
    A = np.identity(n) + BHtRinv @ H
    b1 = xb + BHtRinv @ np.asarray(y, dtype=np.float64).ravel()
    return cg(A, b1, xb, tol=1e-5, maxiter=1000)[0]


def analysis_diagonal(xb, y, loc, weight):
    """Closed form for default H / R / B: weight[r] = valid[r] / sd^2 (0 for masked rows)."""
    xb = np.asarray(xb, dtype=np.float64)
    num = xb.copy()
    den = np.ones_like(xb)
    np.add.at(num, loc, weight * np.asarray(y, dtype=np.float64))
    np.add.at(den, loc, weight)
    return num / den


def cycle(input_state_N, start_time, obs_values, obs_times, obs_system_index, n_cycles, obs_error_sd,
          forecast_fn, delta_t, stationary_observers=True, analysis_window=0.2,
          analysis_time_in_window=None, return_forecast=False, literal=True):
    """DACycler.cycle for Var3D: x[cycle, index] (analyses) or x[cycle, cycle_timestep, index]."""
    x = np.array(input_state_N, dtype=np.float64, copy=True)
    N = x.size
    obs_values = np.asarray(obs_values, dtype=np.float64)
    obs_system_index = np.asarray(obs_system_index, dtype=np.int64)
    if analysis_time_in_window is None:
        analysis_time_in_window = analysis_window / 2
    steps_per_window = round(analysis_window / delta_t) + 1
    time_offset = analysis_window / 2 - analysis_time_in_window
    all_times = _win.get_all_times(start_time, analysis_window, n_cycles)
    filt = _win.get_obs_indices(all_times + time_offset, obs_times, analysis_window,
                                start_inclusive=True, end_inclusive=False)
    padded = _win.pad_time_indices(filt, add_one=True)
    if stationary_observers:
        loc_masks = np.ones(obs_values.shape, dtype=bool)
    else:
        loc_masks = ~np.isnan(obs_values)
        obs_values = np.nan_to_num(obs_values, nan=0.0)
    out = []
    n_obs = obs_values.shape[-1]
    for c in range(n_cycles):
        fidx = padded[c]
        time_mask = fidx > 0
        fidx = fidx - 1
        vals, locs, lmask = obs_values[fidx], obs_system_index[fidx], loc_masks[fidx]
        tmask = np.repeat(time_mask, n_obs)
        if literal:
            H = calc_default_H(vals.size, N, locs)
            R = calc_default_R(vals.size, obs_error_sd)
            H = np.where(tmask, H.T, 0).T                      # _var3d.py:88-89
            H = np.where(lmask.ravel(), H.T, 0).T
            xa = analysis_literal(x, vals, H, R)
        else:
            w = (tmask & lmask.ravel()).astype(np.float64) / (obs_error_sd ** 2)
            xa = analysis_diagonal(x, vals.ravel(), locs.ravel(), w)
        traj = forecast_fn(xa, steps_per_window)               # [S+1, N]
        x = traj[-1]
        out.append(traj)
    out = np.stack(out)
    return out[:, :-1, :] if return_forecast else out[:, 0, :]
