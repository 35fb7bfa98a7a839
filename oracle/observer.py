"""Oracle: observation sampling (TEST INFRASTRUCTURE ONLY).

Restates `dabench.observer.Observer.observe` (dabench/observer/_observer.py:273-366)
for single-variable 1-D states (Lorenz-96: variable 'x', coordinate 'index'), i.e. the
input contract of the ETKF hot path.  The order in which the single
`np.random.default_rng(random_seed)` stream is consumed is what makes the index sets
reproducible (:282, :192-207, :209-234, :236-271, :346-348) and is followed exactly.

Returns a plain dict (the reference returns an xarray Dataset with the same fields):
  time          [n_times]          observation times
  time_index    [n_times]          their indices into the trajectory
  system_index  [n_times, n_obs]   int64 flattened-state indices (0 in padded slots)
  values        [n_times, n_obs]   trajectory value + error (NaN in padded slots)
  errors        [n_times, n_obs]
  stationary_observers, error_sd   (Dataset attrs in the reference, :352-354)
"""
import numpy as np


def observe(traj, traj_times, random_time_density=1.0, random_location_density=1.0,
            random_time_count=None, random_location_count=None, times=None,
            locations=None, stationary_observers=True, error_bias=0.0, error_sd=0.0,
            error_positive_only=False, random_seed=99):
    traj = np.asarray(traj, dtype=np.float64)            # [n_traj_times, N]
    traj_times = np.asarray(traj_times, dtype=np.float64)
    n_t, N = traj.shape
    coord = np.arange(N)
    rng = np.random.default_rng(random_seed)             # :282

    if times is None:                                    # :192-207
        if random_time_count is not None:
            times = np.sort(rng.choice(traj_times, size=random_time_count,
                                       replace=False, shuffle=False))
        else:
            times = traj_times[np.where(
                rng.binomial(1, p=random_time_density, size=n_t).astype(bool))[0]]
    times = np.asarray(times, dtype=np.float64)
    # .sel(time=...) is an exact-label lookup
    tidx = np.searchsorted(traj_times, times)
    if np.any(tidx >= n_t) or np.any(traj_times[np.minimum(tidx, n_t - 1)] != times):
        raise KeyError('observation times must be trajectory times')
    n_times = len(times)

    if stationary_observers:
        if locations is None:                            # :209-234
            if random_location_count is not None:
                count = random_location_count
            else:
                count = int(np.sum(rng.binomial(1, p=random_location_density, size=N)))
            locs = rng.choice(coord, size=count, replace=False, shuffle=False)
        else:
            locs = np.asarray(locations, dtype=np.int64)
            count = len(locs)
        location_dim = count
        sysidx = np.tile(locs, (n_times, 1)).astype(np.int64)
        clean = traj[tidx][:, locs]
    else:
        if locations is None:                            # :236-271
            if random_location_count is not None:
                counts = [random_location_count] * n_times
            else:
                counts = [int(np.sum(rng.binomial(1, p=random_location_density, size=N)))
                          for _ in range(n_times)]
            loc_list = [rng.choice(coord, size=lc, replace=False, shuffle=False)
                        for lc in counts]
        else:
            loc_list = [np.asarray(l, dtype=np.int64) for l in locations]
            counts = [len(l) for l in loc_list]
        location_dim = int(np.max(counts)) if counts else 0
        sysidx = np.zeros((n_times, location_dim), dtype=np.int64)    # NaN -> 0, :327-328
        clean = np.full((n_times, location_dim), np.nan)              # pad -> NaN, :312-324
        for i, l in enumerate(loc_list):
            sysidx[i, :len(l)] = l
            clean[i, :len(l)] = traj[tidx[i], l]

    # one normal draw for every slot, padded ones included (:331-348)
    errors = rng.normal(loc=error_bias, scale=error_sd, size=(1, n_times, location_dim))[0]
    if error_positive_only:
        errors[errors < 0.0] = 0.0
    return dict(time=times, time_index=tidx, system_index=sysidx,
                values=clean + errors, errors=errors,
                stationary_observers=stationary_observers, error_sd=error_sd)
