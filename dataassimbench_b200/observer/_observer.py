"""Observation sampling: the input contract of `ETKF.cycle()`.

Mirrors `dabench.observer.Observer` (dabench/observer/_observer.py:89-366) for single-variable
states: 1-D (`x[time, index]`, Lorenz-96 / -63) or gridded (`pv[time, level, x, y]`, SQGTurb: one location draw
per coordinate, with replacement, :223-234).  Index sets must be bit-compatible with the reference, so the one
`numpy.random.default_rng(random_seed)` stream is consumed in the reference's order: times
(:192-207), then locations (:209-271), then ONE normal draw covering every slot (:346-348).
The draws are host-side NumPy.  The sampling itself -- `state_vec.sel(time=...).sel(locations)`,
pad, `+ errors` (:290-316, 361-363) -- runs on the device (`dab_observe_f64`) when the trajectory
is a CUDA tensor, e.g. a nature run produced by `dab_l96_forecast_f64`: only the sampled values
leave HBM.  With a NumPy trajectory it is plain indexing.
"""
import numpy as np

from .. import _xr


class Observer:
    def __init__(self, state_vec, random_time_density=1., random_location_density=1.,
                 random_time_count=None, random_location_count=None, times=None, locations=None,
                 stationary_observers=True, error_bias=0., error_sd=0., error_positive_only=False,
                 random_seed=99, store_as_jax=False, keep_on_device=False):
        self.state_vec = state_vec
        # not a reference argument: with a trajectory in HBM, leave the sampled values there too (the 'x' variable
        # of the result is then a CUDA tensor, which ETKF.cycle() packs into its tables without a host round trip)
        self.keep_on_device = keep_on_device
        self.random_time_density = random_time_density
        self.random_location_density = random_location_density
        self.random_time_count = random_time_count
        self.random_location_count = random_location_count
        self.times = None if times is None else np.array(times)
        self.locations = locations
        self.stationary_observers = stationary_observers
        self.error_bias = error_bias
        self.error_sd = error_sd
        self.error_positive_only = error_positive_only
        self.random_seed = random_seed
        self.store_as_jax = store_as_jax
        if np.ndim(error_bias) != 0 or np.ndim(error_sd) != 0:
            raise NotImplementedError('per-location error_bias / error_sd lists are outside the ETKF path')

    def _draw_count(self, rng, n):
        if self.random_location_count is not None:
            return int(self.random_location_count)
        return int(np.sum(rng.binomial(1, p=self.random_location_density, size=n)))

    def _draw_locations(self, rng, coords, count):
        """One `choice` per non-time coordinate, in the Dataset's order; with more than one coordinate the draws
        are WITH replacement (dabench/observer/_observer.py:223-234, :258-269).  Returns flattened-state indices."""
        replace = len(coords) > 1
        picks = [rng.choice(c, size=count, replace=replace, shuffle=False) for c in coords]
        if len(coords) == 1:
            return np.asarray(picks[0], dtype=np.int64)
        return np.ravel_multi_index(tuple(np.asarray(q, dtype=np.int64) for q in picks),
                                    tuple(len(c) for c in coords)).astype(np.int64)

    def observe(self):
        # single-variable states: `x[time, index]` (Lorenz-96 / -63) or any `var[time, c1, c2, ...]` such as the
        # SQGTurb `pv[time, level, x, y]`; locations are drawn per coordinate and sampled from the flattened state
        names = [str(v) for v in self.state_vec.data_vars]
        var = 'x' if 'x' in names else names[0]
        self._var = var
        dims = tuple(_xr.var_dims(self.state_vec, var))
        if not dims or dims[0] != 'time':
            raise ValueError("the state variable must have 'time' as its first dimension")
        traj = _xr.var_data(self.state_vec, var)
        on_device = _is_cuda_tensor(traj)
        if not on_device:
            traj = _xr.var_array(self.state_vec, var)
        traj_t = _xr.var_array(self.state_vec, 'time')
        grid_shape = tuple(traj.shape[1:])
        traj = traj.reshape(traj.shape[0], -1)
        n_t, n_sys = traj.shape
        coords = [np.arange(n) for n in grid_shape]           # positional coordinates, as Data.generate builds them
        rng = np.random.default_rng(self.random_seed)

        if self.times is None:
            if self.random_time_count is not None:
                self.times = np.sort(rng.choice(traj_t, size=self.random_time_count,
                                                replace=False, shuffle=False))
            else:
                keep = rng.binomial(1, p=self.random_time_density, size=n_t).astype(bool)
                self.times = traj_t[np.where(keep)[0]]
        self.time_dim = self.times.shape[0]
        pos = np.searchsorted(traj_t, self.times)
        if np.any(pos >= n_t) or np.any(traj_t[np.minimum(pos, n_t - 1)] != self.times):
            raise KeyError('observation times must be times of the trajectory')

        if self.stationary_observers:
            if self.locations is None:
                count = self._draw_count(rng, n_sys)
                locs = self._draw_locations(rng, coords, count)
            else:
                locs = np.asarray(_loc_array(self.locations), dtype=np.int64)
            self.location_dim = len(locs)
            sysidx = np.tile(locs, (self.time_dim, 1)).astype(np.int64)
            counts = None
        else:
            if self.locations is None:
                if self.random_location_count is not None:
                    counts = [int(self.random_location_count)] * self.time_dim
                else:       # every count first, then every location draw (:246-268)
                    counts = [self._draw_count(rng, n_sys) for _ in range(self.time_dim)]
                per_time = [self._draw_locations(rng, coords, c) for c in counts]
            else:
                per_time = [np.asarray(_loc_array(l), dtype=np.int64) for l in self.locations]
                counts = [len(l) for l in per_time]
            self._location_counts = counts
            self.location_dim = int(np.max(counts)) if len(counts) else 0
            sysidx = np.zeros((self.time_dim, self.location_dim), dtype=np.int64)
            for i, l in enumerate(per_time):
                sysidx[i, :len(l)] = l

        errors = rng.normal(loc=self.error_bias, scale=self.error_sd,
                            size=(1, self.time_dim, self.location_dim))
        if self.error_positive_only:
            errors[errors < 0.] = 0.
        if on_device:
            values = self._sample_on_device(traj, pos, sysidx, counts, errors[0])
        else:
            if counts is None:
                clean = traj[pos][:, sysidx[0]] if self.time_dim else np.empty((0, self.location_dim))
            else:
                clean = np.full((self.time_dim, self.location_dim), np.nan)
                for i, c in enumerate(counts):
                    clean[i, :c] = traj[pos[i], sysidx[i, :c]]
            values = clean + errors[0]
        return _xr.new_dataset(
            {var: (('time', 'observations'), values),
             'system_index': (('variable', 'time', 'observations'), sysidx[None]),
             'errors': (('variable', 'time', 'observations'), errors)},
            coords={'time': self.times, 'variable': np.array([var])},
            attrs={'stationary_observers': self.stationary_observers, 'error_sd': self.error_sd})

    def _sample_on_device(self, traj, pos, sysidx, counts, errors):
        """values[t][j] = traj[pos[t]][sysidx[t][j]] + errors[t][j], NaN beyond counts[t]: one
        `dab_observe_f64` launch on the trajectory's device; the index tables and the errors go up,
        only the sampled values come back."""
        import torch
        from .._runtime import handle
        if traj.dtype != torch.float64 or not traj.is_contiguous():
            raise TypeError("a device trajectory must be a contiguous float64 tensor [time, index]")
        dev = traj.device
        n_t, n_obs = sysidx.shape
        if n_t * n_obs == 0:
            return np.empty((n_t, n_obs))
        H = handle(dev.index if dev.index is not None else torch.cuda.current_device())
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        stationary = counts is None
        d_pos = up(np.asarray(pos, dtype=np.int64))
        d_loc = up(sysidx[0] if stationary else sysidx)
        d_cnt = None if stationary else up(np.asarray(counts, dtype=np.int64))
        d_err = up(np.asarray(errors, dtype=np.float64))
        d_val = torch.empty((n_t, n_obs), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream()
            H.observe(traj, traj.shape[1], traj.shape[0], d_pos, n_t, d_loc, stationary, d_cnt, n_obs,
                      d_err, d_val, stream=stream.cuda_stream)
        return d_val if self.keep_on_device else d_val.cpu().numpy()


def _is_cuda_tensor(a):
    return type(a).__module__.split('.')[0] == 'torch' and bool(getattr(a, 'is_cuda', False))


def _loc_array(locations):
    """Accepts an index array, or the reference's {'index': DataArray} mapping."""
    if isinstance(locations, dict):
        locations = next(iter(locations.values()))
    return np.asarray(locations.values if hasattr(locations, 'values') else locations)
