"""Base cycler: constructor contract and the host side of `cycle()`.

Mirrors `dabench.dacycler.DACycler` (dabench/dacycler/_dacycler.py:18-56, 186-290).  The
per-cycle body the reference traces with `jax.lax.scan` (`_cycle_and_forecast`, :110-141) is
replaced by the device-resident loop of the C ABI (`dab_cycler_*`); this class only derives the
window tables and marshals arrays.
"""
import numpy as np

from .. import _xr
from . import _windows


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
