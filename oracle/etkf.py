"""Oracle: ETKF analysis and the cycling driver (TEST INFRASTRUCTURE ONLY).

Restates, with NumPy/SciPy in FP64:
  * default H / R builders     -- dabench/dacycler/_dacycler.py:58-72
  * mask application           -- dabench/dacycler/_etkf.py:202-203
  * ETKF._compute_analysis     -- dabench/dacycler/_etkf.py:119-163  (`analysis_literal`)
  * DACycler._cycle_and_forecast / cycle()
                               -- dabench/dacycler/_dacycler.py:110-141, 186-290  (`cycle`)

`analysis_sparse` is the algebraically equivalent formulation the CUDA kernels
implement (index gather for H, diagonal R, symmetric eigendecomposition, one
k x k transform); tests/test_oracle_golden.py proves it agrees with the literal
formula to ~1e-15 (max-norm relative) wherever the literal formula can be held in memory.

Layout convention of this module: ensembles are [k, N] (member-major), i.e. the
`x[ensemble, index]` array of the reference's input Dataset; the reference's
internal Xb[N, k] is its transpose (dabench/dacycler/_etkf.py:195).
"""
import numpy as np
import scipy.linalg

from . import windows as _win


# --------------------------------------------------------------------------- operators
def calc_default_H(n_y, system_dim, obs_loc_indices):
    """dabench/dacycler/_dacycler.py:58-66 -- dense one-hot H[n_y, N]."""
    H = np.zeros((n_y, system_dim))
    H[np.arange(n_y), np.asarray(obs_loc_indices).ravel()] = 1
    return H


def calc_default_R(n_y, obs_error_sd):
    """dabench/dacycler/_dacycler.py:68-72."""
    return np.identity(n_y) * (obs_error_sd ** 2)


# --------------------------------------------------------------------------- analysis
def analysis_literal(Xb_Nk, Y, H, R, rho=1.0):
    """Line-by-line restatement of ETKF._compute_analysis (dabench/dacycler/_etkf.py:119-163).

    Xb_Nk: [N, k]; Y: obs values (any shape, flattened); H: [n_y, N]; R: [n_y, n_y].
    jnp.linalg.pinv(rtol=1e-15) -> np.linalg.pinv(rcond=1e-15) (same cutoff rule);
    jax.scipy.linalg.sqrtm(..).real -> scipy.linalg.sqrtm(..).real (same Schur method).
    """
    system_dim, ensemble_dim = Xb_Nk.shape
    U = np.ones((ensemble_dim, ensemble_dim)) / ensemble_dim
    I = np.identity(ensemble_dim)

    Xb_pert = Xb_Nk @ (I - U)
    Xb = Xb_pert + Xb_Nk @ U

    Yb = H @ Xb

    Xb_bar = np.mean(Xb, axis=1)
    Xb_pert = Xb @ (I - U)
    yb_bar = np.mean(Yb, axis=1)
    Yb_pert = Yb @ (I - U)

    if len(R) > 0:
        Rinv = np.linalg.pinv(R, rcond=1e-15)
        Pa_ens = np.linalg.pinv((ensemble_dim - 1) / rho * I + Yb_pert.T @ Rinv @ Yb_pert,
                                rcond=1e-15)
        Wa = scipy.linalg.sqrtm((ensemble_dim - 1) * Pa_ens)
        Wa = np.real(Wa)
    else:
        Rinv = np.zeros_like(R)
        Pa_ens = np.zeros((ensemble_dim, ensemble_dim))
        Wa = np.zeros((ensemble_dim, ensemble_dim))

    wa = Pa_ens @ Yb_pert.T @ Rinv @ (np.asarray(Y).ravel() - yb_bar)
    Xa_pert = Xb_pert @ Wa
    Xa_bar = Xb_bar + np.ravel(Xb_pert @ wa)
    v = np.ones((1, ensemble_dim))
    Xa = Xa_pert + Xa_bar[:, None] @ v
    return Xa


def obs_gather(X_kN, idx):
    """Yb[j, r] = X[j, idx[r]] -- what `H @ Xb` computes for a one-hot H (bit-exact copy)."""
    return np.asarray(X_kN)[:, np.asarray(idx, dtype=np.int64)]


def contraction(X_kN, y, idx, w):
    """C = Y'^T R^-1 Y' (k x k) and g = Y'^T R^-1 (y - ybar) (k) with diagonal R^-1 = w
    (w = 0 for masked rows).  dabench/dacycler/_etkf.py:137-138,144-146,154."""
    Yb = obs_gather(X_kN, idx)
    ybar = Yb.mean(axis=0)
    Yp = Yb - ybar
    Yw = Yp * w
    return Yw @ Yp.T, Yw @ (np.asarray(y, dtype=np.float64) - ybar)


def transform_from_stats(C, g, k, rho, empty_R=False):
    """k x k matrix T with Xa[N,k] = Xb[N,k] @ T  (T = U + (I-U)(wa 1^T + Wa)).

    Eigendecomposition form of dabench/dacycler/_etkf.py:142-161:
      A = (k-1)/rho I + C = V L V^T, Pa = V L^-1 V^T (pinv cutoff 1e-15*max L),
      Wa = sqrtm((k-1) Pa) = V sqrt((k-1)/L) V^T, wa = Pa g.
    """
    U = np.ones((k, k)) / k
    I = np.identity(k)
    if empty_R:   # dabench/dacycler/_etkf.py:149-152
        return U.copy()
    A = (k - 1) / rho * I + C
    lam, V = np.linalg.eigh(A)
    linv = np.where(lam > 1e-15 * lam.max(), 1.0 / lam, 0.0)
    Pa = (V * linv) @ V.T
    Wa = (V * np.sqrt((k - 1) * linv)) @ V.T
    wa = Pa @ g
    return U + (I - U) @ (wa[:, None] @ np.ones((1, k)) + Wa)


def analysis_sparse(X_kN, y, idx, w, rho=1.0, empty_R=False):
    """Equivalent sparse analysis on the [k, N] layout; returns Xa [k, N]."""
    X_kN = np.asarray(X_kN, dtype=np.float64)
    k = X_kN.shape[0]
    if empty_R:
        T = transform_from_stats(None, None, k, rho, empty_R=True)
    else:
        C, g = contraction(X_kN, y, idx, w)
        T = transform_from_stats(C, g, k, rho)
    return T.T @ X_kN


# --------------------------------------------------------------------------- cycling
def cycle(input_state_kN, start_time, obs_values, obs_times, obs_system_index,
          n_cycles, obs_error_sd, forecast_fn, delta_t,
          stationary_observers=True, analysis_window=0.2, analysis_time_in_window=None,
          return_forecast=False, rho=1.0, literal=True):
    """Restatement of DACycler.cycle for the ETKF (dabench/dacycler/_dacycler.py:186-290).

    input_state_kN : [k, N]       initial ensemble (`x[ensemble, index]`)
    obs_values     : [n_times, n_obs] (NaN padded when nonstationary)
    obs_system_index: [n_times, n_obs] int
    forecast_fn(x[N], n_points) -> [n_points, N] trajectory incl. the initial state
                     (the `Model.forecast` plugin; tests/dacycler_etkf_test.py:39-46)
    delta_t        : the *cycler's* delta_t (only sets steps_per_window, :233)
    Returns x[cycle, ensemble, cycle_timestep, index] (return_forecast) or
            x[cycle, ensemble, index] (analyses only)              (:281-290).
    """
    X = np.array(input_state_kN, dtype=np.float64, copy=True)
    k, N = X.shape
    obs_values = np.asarray(obs_values, dtype=np.float64)
    obs_system_index = np.asarray(obs_system_index, dtype=np.int64)

    if analysis_time_in_window is None:
        analysis_time_in_window = analysis_window / 2          # :226-230 (3D branch)
    steps_per_window = round(analysis_window / delta_t) + 1    # :233
    time_offset = analysis_window / 2 - analysis_time_in_window  # :237

    all_times = _win.get_all_times(start_time, analysis_window, n_cycles)
    filt = _win.get_obs_indices(all_times + time_offset, obs_times, analysis_window,
                                start_inclusive=True, end_inclusive=False)
    padded = _win.pad_time_indices(filt, add_one=True)         # :259

    if stationary_observers:                                   # :262-268
        loc_masks = np.ones(obs_values.shape, dtype=bool)
    else:
        loc_masks = ~np.isnan(obs_values)
        obs_values = np.nan_to_num(obs_values, nan=0.0)

    out = []
    n_obs = obs_values.shape[-1]
    for c in range(n_cycles):
        fidx = padded[c]
        time_mask = fidx > 0                                   # :119
        fidx = fidx - 1                                        # :120 (negative wraps, masked)
        if len(fidx) and obs_values.shape[0]:
            vals = obs_values[fidx]                            # [T_pad, n_obs]
            locs = obs_system_index[fidx]
            lmask = loc_masks[fidx]
        else:
            vals = np.zeros((len(fidx), n_obs))
            locs = np.zeros((len(fidx), n_obs), dtype=np.int64)
            lmask = np.zeros((len(fidx), n_obs), dtype=bool)
        tmask = np.repeat(time_mask, n_obs)                    # :126
        n_y = vals.size
        if literal:
            H = calc_default_H(n_y, N, locs)
            R = calc_default_R(n_y, obs_error_sd)
            H = np.where(tmask, H.T, 0).T                      # _etkf.py:202
            H = np.where(lmask.ravel(), H.T, 0).T              # _etkf.py:203
            Xa = analysis_literal(X.T, vals, H, R, rho).T
        else:
            w = (tmask & lmask.ravel()).astype(np.float64) / (obs_error_sd ** 2)
            Xa = analysis_sparse(X, vals.ravel(), locs.ravel(), w, rho, empty_R=(n_y == 0))
        # forecast each member over the window (_etkf.py:64-80)
        traj = np.stack([forecast_fn(Xa[i], steps_per_window) for i in range(k)])  # [k, S+1, N]
        X = traj[:, -1, :]
        out.append(traj)
    out = np.stack(out)                                        # [cycle, ensemble, time, index]
    if return_forecast:
        return out[:, :, :-1, :]                               # drop_isel(cycle_timestep=-1)
    return out[:, :, 0, :]                                     # isel(cycle_timestep=0)
