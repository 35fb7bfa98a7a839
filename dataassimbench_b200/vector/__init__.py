"""Legacy array containers with the behaviour of `dabench.vector` (dabench/vector/_vector.py:9-151,
_state_vector.py:8-99, _obs_vector.py:12-185): time-indexed `values` with `times`, slicing / integer /
boolean-mask selection that returns an independent copy, `filter_times`, the `xi` ("latest state") and
gridded views of a `StateVector`, `obs_dims` bookkeeping and the per-observation side arrays of an
`ObsVector`.

`ETKF.cycle()` itself works on Datasets; these classes are what the
reference's older call sites pass around (`ObsOp.h(state_vec)`, `Observer` inputs), so they keep
the reference's attribute names, defaults and error messages.  JAX is not part of this package:
`store_as_jax` is accepted and remembered, the arrays are NumPy either way.
"""
import copy
import warnings

import numpy as np

__all__ = ['StateVector', 'ObsVector']

_NO_DATA = ('Object does not contain any data values.\n'
            'Run .generate() or .load() and try again')


def _array_property(name):
    """A property that stores `np.asarray(value)` (or None) in `_<name>`."""
    private = '_' + name

    def getter(self):
        return getattr(self, private)

    def setter(self, vals):
        setattr(self, private, None if vals is None else np.asarray(vals))

    def deleter(self):
        delattr(self, private)

    return property(getter, setter, deleter)


class _Vector:
    """`values[time, ...]` with `times[time]` (dabench/vector/_vector.py:9-151)."""

    # arrays indexed along the time axis together with `values` by __getitem__
    _time_indexed = ('values', 'times')

    times = _array_property('times')

    def __init__(self, time_dim=None, times=None, store_as_jax=False, values=None):
        self.store_as_jax = store_as_jax
        self.values = values
        # dabench/vector/_vector.py:28-35: default times 0..T-1, time_dim from times
        self.times = np.arange(self.values.shape[0]) if (times is None and self.values is not None) else times
        self.time_dim = self.times.shape[0] if (time_dim is None and self.times is not None) else time_dim

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, vals):
        self._values = None if vals is None else np.asarray(vals)

    @values.deleter
    def values(self):
        del self._values

    def _select(self, subscript):
        """Deep copy with every time-indexed array cut by `subscript` (slice, integer, index or mask array)."""
        if self.values is None:
            raise AttributeError(_NO_DATA)
        out = copy.deepcopy(self)
        if not isinstance(subscript, (slice, int, np.integer)):
            subscript = np.asarray(subscript)
            if subscript.dtype != bool and not np.issubdtype(subscript.dtype, np.integer):
                # the reference's filter_times hands over a 0/1 float product of conditions
                subscript = subscript.astype(bool)
        for name in self._time_indexed:
            cur = getattr(out, name, None)
            if cur is not None:
                setattr(out, name, cur[subscript])
        return out

    def __getitem__(self, subscript):
        out = self._select(subscript)
        # dabench/vector/_vector.py:42-59: a single integer leaves one state
        out.time_dim = 1 if isinstance(subscript, (int, np.integer)) else out.times.shape[0]
        return out

    def filter_times(self, start, end, inclusive=True, start_inclusive=None, end_inclusive=None):
        """Copy restricted to start <= t <= end (dabench/vector/_vector.py:97-151).  Equality with an end
        point is `isclose(rtol=0)` for floats and exact for datetimes; naming one of start_inclusive /
        end_inclusive switches the other one off unless it is given too."""
        if start_inclusive is not None or end_inclusive is not None:
            start_inclusive = bool(start_inclusive)
            end_inclusive = bool(end_inclusive)
        elif inclusive:
            start_inclusive = end_inclusive = True
        else:
            start_inclusive = end_inclusive = False

        def at(point):
            if isinstance(point, np.datetime64):
                return self.times == point
            return np.isclose(self.times, point, rtol=0)

        keep = np.ones(self.times.shape[0], dtype=bool)
        if start is not None:
            eq = at(start)
            keep &= ((self.times > start) | eq) if start_inclusive else ((self.times > start) & ~eq)
        if end is not None:
            eq = at(end)
            keep &= ((self.times < end) | eq) if end_inclusive else ((self.times < end) & ~eq)
        return self[keep]


class StateVector(_Vector):
    """State(s) of a system, `values[time_dim, system_dim]` (dabench/vector/_state_vector.py:8-99)."""

    def __init__(self, system_dim=None, time_dim=None, original_dim=None, delta_t=None, store_as_jax=False,
                 values=None, times=None, **kwargs):
        self._xi = None
        self.system_dim = system_dim
        self.delta_t = delta_t
        self.original_dim = system_dim if original_dim is None else original_dim
        super().__init__(time_dim=time_dim, store_as_jax=store_as_jax, values=values, times=times, **kwargs)

    def __str__(self):
        return f'Current State = {self.xi}, Timesteps = {self.time_dim}'

    @property
    def xi(self):
        """Most recent state; defaults to the last row of `values`."""
        if self._xi is None and self.values is not None:
            self.xi = self.values[-1]
        return self._xi

    @xi.setter
    def xi(self, vals):
        self._xi = None if vals is None else np.asarray(vals)

    @xi.deleter
    def xi(self):
        del self._xi

    def _original_shape(self):
        od = self.original_dim
        return tuple(od) if np.ndim(od) else (int(od),)

    @property
    def xi_gridded(self):
        return None if self._xi is None else self._xi.reshape(self._original_shape())

    @property
    def values_gridded(self):
        """`values` back in the generator's grid: (time_dim, *original_dim)."""
        return None if self._values is None else self._to_original_dim()

    def _to_original_dim(self):
        return self.values.reshape((self.time_dim,) + self._original_shape())


class ObsVector(_Vector):
    """Observations (dabench/vector/_obs_vector.py:12-185): `values[num_obs(, obs_dim)]` with per-observation
    `times`, `location_indices`, `errors`, `coords`."""

    _time_indexed = ('values', 'times', 'location_indices', 'errors', 'coords')

    coords = _array_property('coords')
    errors = _array_property('errors')

    def __init__(self, num_obs=None, obs_dims=None, error_dist=None, values=None, coords=None,
                 time_indices=None, location_indices=None, errors=None, error_sd=None, error_bias=None,
                 times=None, store_as_jax=False, stationary_observers=True, **kwargs):
        self.num_obs = num_obs
        self.obs_dims = None
        self.error_dist = error_dist
        self.error_sd = error_sd
        self.error_bias = error_bias
        self.time_indices = time_indices
        self.location_indices = location_indices
        self.stationary_observers = stationary_observers
        self._coords = self._errors = None
        super().__init__(times=times, store_as_jax=store_as_jax, **kwargs)
        self.values = values
        if self.values is None:
            self.obs_dims = obs_dims
        elif obs_dims is not None and not np.array_equal(obs_dims, self.obs_dims):
            warnings.warn('obs_dims {} does not match dimensions of '
                          ' values {}.\n Proceeding with obs_dims '
                          'calculated based on values..'.format(obs_dims, self.obs_dims))
        self.coords = coords
        self.errors = errors

    def __getitem__(self, subscript):
        # dabench/vector/_obs_vector.py:91-117: the side arrays follow, num_obs / obs_dims through the setter
        return self._select(subscript)

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, vals):
        if vals is None:
            self._values = None
            return
        self._values = np.asarray(vals)
        if self._values.ndim == 0:                 # a single scalar observation picked by an integer index
            self.num_obs, self.obs_dims = 1, np.array([1])
            return
        self.num_obs = self._values.shape[0]
        self._calc_obs_dims()

    @values.deleter
    def values(self):
        del self._values

    def _calc_obs_dims(self):
        """Values per observation: ragged object arrays, scalars, or a fixed second dimension."""
        v = self._values
        if v.dtype == np.dtype('O'):
            self.obs_dims = np.array([np.shape(x)[0] for x in v])
        elif v.ndim == 1:
            self.obs_dims = np.repeat(1, v.shape[0])
        else:
            self.obs_dims = np.repeat(v.shape[1], v.shape[0])
