"""Base cycler: constructor contract and the host side of `cycle()`.

Mirrors `dabench.dacycler.DACycler` (dabench/dacycler/_dacycler.py:18-56, 186-290).  The
per-cycle body the reference traces with `jax.lax.scan` (`_cycle_and_forecast`, :110-141) is
replaced by the device-resident loop of the C ABI (`dab_cycler_*`); this class only derives the
window tables and marshals arrays.
"""
import numpy as np

from .. import _xr
from . import _windows


class IndexOperator:
    """The default observation operator without its dense form.

    `DACycler._calc_default_H` (dabench/dacycler/_dacycler.py:58-66) builds a dense one-hot matrix
    `H[r, loc[r]] = 1` (10 GB at N=65536 with 30 % observed); the kernels take the index list itself
    (`dab_obs_gather_f64`: `(H @ Xb)[r] = Xb[loc[r]]`), so that is what this object carries, together
    with the row mask the reference applies by zeroing rows of H (dabench/dacycler/_etkf.py:202-203).
    `dense()` materialises the reference's matrix for small cases and tests."""

    def __init__(self, indices, system_dim, mask=None):
        self.indices = np.ascontiguousarray(np.ravel(indices), dtype=np.int64)
        self.system_dim = int(system_dim)
        self.mask = None if mask is None else np.ascontiguousarray(np.ravel(mask)).astype(bool)
        if self.mask is not None and self.mask.shape != self.indices.shape:
            raise ValueError('mask and indices differ in length')

    @property
    def shape(self):
        return (self.indices.shape[0], self.system_dim)

    def masked(self, mask):
        m = np.ravel(np.asarray(mask)).astype(bool)
        return IndexOperator(self.indices, self.system_dim, m if self.mask is None else (self.mask & m))

    def dense(self):
        H = np.zeros(self.shape)
        rows = np.arange(self.indices.shape[0])
        keep = np.ones_like(rows, dtype=bool) if self.mask is None else self.mask
        H[rows[keep], self.indices[keep]] = 1.0
        return H

    @staticmethod
    def from_dense(H):
        """A dense matrix whose rows are one-hot (or zero: masked) as an IndexOperator; anything else is
        outside the accelerated contract."""
        H = np.asarray(H)
        if H.ndim != 2:
            raise NotImplementedError('H must be an IndexOperator, an index array or a dense one-hot matrix')
        nz = H != 0
        cnt = nz.sum(axis=1)
        if (cnt > 1).any() or not np.all(H[nz] == 1.0):
            raise NotImplementedError('only one-hot (index-gather) observation operators are accelerated; '
                                      'this dense H mixes grid points')
        return IndexOperator(np.where(cnt == 1, nz.argmax(axis=1), 0), H.shape[1], cnt == 1)


class ScaledIdentity:
    """`sd**2 * I` without the dense matrix (`DACycler._calc_default_R`, dabench/dacycler/_dacycler.py:68-72;
    3 GB at 19 661 observations).  `len()` is what the reference's empty-R test looks at
    (dabench/dacycler/_etkf.py:141)."""

    def __init__(self, n, sd):
        self.n = int(n)
        self.sd = float(sd)

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return (self.n, self.n)

    def dense(self):
        return np.identity(self.n) * self.sd ** 2

    @staticmethod
    def sd_of(R):
        """sigma of a DIAGONAL R: a float when R = sd**2 * I (ScaledIdentity, or a dense matrix of that form),
        a vector of per-observation sigmas when the diagonal entries differ (a 1-D array is taken as the
        diagonal of variances).  Anything with off-diagonal entries raises NotImplementedError."""
        if isinstance(R, ScaledIdentity):
            return R.sd
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 1:
            d = R
        elif R.ndim == 2 and R.shape[0] == R.shape[1]:
            d = np.diag(R)
            if np.count_nonzero(R - np.diag(d)) != 0:
                raise NotImplementedError('only a diagonal R is accelerated')
        else:
            raise NotImplementedError('only a diagonal R is accelerated')
        if d.shape[0] == 0:
            return 1.0
        if np.any(d < 0):
            raise NotImplementedError('only a diagonal R with non-negative variances is accelerated')
        if np.all(d == d[0]):
            return float(np.sqrt(d[0]))
        return np.sqrt(d)


class DACycler:
    _in_4d = False
    _uses_ensemble = False

    def __init__(self, system_dim, delta_t, model_obj, B=None, R=None, H=None, h=None):
        self.h = h
        self.H = H
        self.R = R
        self.B = B
        self.system_dim = system_dim
        self.delta_t = delta_t
        self.model_obj = model_obj

    # ---- host bookkeeping shared by every cycler (dabench/dacycler/_dacycler.py:217-268)
    def _prepare_cycle(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd,
                       analysis_window, analysis_time_in_window):
        self._observed_vars = list(np.atleast_1d(_xr.var_array(obs_vector, 'variable')))
        self._data_vars = list(input_state.data_vars)
        if obs_error_sd is None:
            obs_error_sd = _xr.get_attr(obs_vector, 'error_sd')
        self.obs_error_sd = obs_error_sd
        self.analysis_window = analysis_window
        if analysis_time_in_window is None:
            analysis_time_in_window = 0 if self._in_4d else analysis_window / 2
        self.steps_per_window = round(analysis_window / self.delta_t) + 1
        time_offset = analysis_window / 2 - analysis_time_in_window
        centres = _windows.window_centres(start_time, analysis_window, n_cycles)
        obs_times = _xr.var_array(obs_vector, 'time')
        padded = _windows.padded_time_table(obs_times, centres + time_offset, analysis_window,
                                            start_inclusive=True, end_inclusive=self._in_4d)
        return dict(centres=centres, padded=padded, obs_times=obs_times,
                    stationary=bool(_xr.get_attr(obs_vector, 'stationary_observers', True)))

    # ---- default operators (dabench/dacycler/_dacycler.py:58-76), in their compact forms
    def _calc_default_H(self, obs_values, obs_loc_indices):
        return IndexOperator(obs_loc_indices, self.system_dim)

    def _calc_default_R(self, obs_values, obs_error_sd):
        return ScaledIdentity(np.size(obs_values), obs_error_sd)

    def _calc_default_B(self):
        return ScaledIdentity(self.system_dim, 1.0)

    def _lorenz96_generator(self):
        """The Lorenz-96 generator behind `model_obj`, or None.  The forecast plugin is recognised by
        what it wraps: the native `Lorenz96Model`, or any `Model` subclass whose `model_obj` is a
        `data.Lorenz96` -- e.g. the wrapper the reference's cycler tests define inline
        (tests/dacycler_etkf_test.py:39-46: `forecast` = `model_obj.generate(x0, n_steps)`).  Its
        `forecast()` body is then not called per member: the forecast kernel integrates that generator
        (its delta_t and forcing_term) for the whole ensemble."""
        from ..data import Lorenz96
        from ..model import Model
        gen = getattr(self.model_obj, 'model_obj', None)
        if isinstance(self.model_obj, Model) and isinstance(gen, Lorenz96):
            return gen
        return None

    def cycle(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd=None,
              analysis_window=0.2, analysis_time_in_window=None, return_forecast=False):
        raise NotImplementedError('only the ETKF cycler is accelerated (dataassimbench_b200.dacycler.ETKF)')
