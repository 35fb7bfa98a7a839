"""Legacy array containers (`dabench.vector`, dabench/vector/_vector.py:9-151,
_state_vector.py:8-99, _obs_vector.py:12-185).  `ETKF.cycle()` does not use them (xarray
Datasets are the real containers); they are kept import-compatible
and hold plain NumPy arrays."""
import numpy as np

__all__ = ['StateVector', 'ObsVector']


class _Vector:
    def __init__(self, times=None, store_as_jax=False, system_dim=None, delta_t=None, values=None, **kwargs):
        self.store_as_jax = store_as_jax
        self.system_dim = system_dim
        self.delta_t = delta_t
        self.times = None if times is None else np.asarray(times)
        self.values = None if values is None else np.asarray(values)
        self.time_dim = None if self.values is None else self.values.shape[0]


class StateVector(_Vector):
    def __init__(self, system_dim=None, time_dim=None, original_dim=None, delta_t=None, values=None,
                 times=None, store_as_jax=False, **kwargs):
        super().__init__(times=times, store_as_jax=store_as_jax, system_dim=system_dim, delta_t=delta_t,
                         values=values)
        if self.values is not None and system_dim is None:
            self.system_dim = int(np.prod(self.values.shape[1:])) if self.values.ndim > 1 \
                else self.values.shape[0]
        self.original_dim = original_dim if original_dim is not None else (self.system_dim,)

    @property
    def values_gridded(self):
        return None if self.values is None else self.values.reshape((-1,) + tuple(self.original_dim))


class ObsVector(_Vector):
    def __init__(self, num_obs=None, obs_dims=None, values=None, locations=None, location_indices=None,
                 errors=None, error_dist=None, error_sd=None, times=None, store_as_jax=False, **kwargs):
        super().__init__(times=times, store_as_jax=store_as_jax, values=values)
        self.locations = locations
        self.location_indices = None if location_indices is None else np.asarray(location_indices)
        self.errors = errors
        self.error_dist = error_dist
        self.error_sd = error_sd
        self.num_obs = num_obs if num_obs is not None else (0 if self.values is None else self.values.shape[0])
        self.obs_dims = obs_dims
