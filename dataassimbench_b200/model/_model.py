"""Forecast-model plugin point (mirrors dabench/model/_model.py:10-32).

`Model` keeps the reference's contract: a subclass must define
`forecast(state_vec, n_steps) -> (last_state, trajectory)` or the constructor raises
`ValueError`.  `Lorenz96Model` is the native implementation the accelerated
`ETKF.cycle()` recognises: the same wrapper the reference's own ETKF test defines inline
(tests/dacycler_etkf_test.py:39-46), backed by the CUDA forecast kernel.
"""
from .. import _xr


class Model:
    def __init__(self, system_dim=None, delta_t=None, model_obj=None):
        self.system_dim = system_dim
        self.delta_t = delta_t
        self.model_obj = model_obj
        if not callable(getattr(self, 'forecast', None)):
            raise ValueError('Model object does not have a defined forecast() method.')


class Lorenz96Model(Model):
    """Wraps a `data.Lorenz96` generator; `forecast` integrates one member for n_steps points."""

    def __init__(self, system_dim=None, delta_t=None, model_obj=None):
        if model_obj is None:
            from ..data import Lorenz96
            model_obj = Lorenz96(system_dim=system_dim or 36, delta_t=delta_t or 0.05)
        super().__init__(system_dim=system_dim or model_obj.system_dim,
                         delta_t=delta_t or model_obj.delta_t, model_obj=model_obj)

    def forecast(self, state_vec, n_steps):
        new_vec = self.model_obj.generate(x0=_xr.var_array(state_vec, 'x'), n_steps=n_steps)
        return new_vec.isel(time=-1), new_vec


class Lorenz63Model(Model):
    """Wraps a `data.Lorenz63` generator the same way (the wrapper the reference's model test defines
    inline, tests/model_base_test.py:10-16, with the n_steps argument of the cyclers' contract)."""

    def __init__(self, system_dim=None, delta_t=None, model_obj=None):
        if model_obj is None:
            from ..data import Lorenz63
            model_obj = Lorenz63(delta_t=delta_t or 0.01)
        super().__init__(system_dim=3, delta_t=delta_t or model_obj.delta_t, model_obj=model_obj)

    def forecast(self, state_vec, n_steps=2):
        new_vec = self.model_obj.generate(x0=_xr.var_array(state_vec, 'x'), n_steps=n_steps)
        return new_vec.isel(time=-1), new_vec


class SQGTurbModel(Model):
    """Wraps a `data.SQGTurb` generator behind the same plugin point: `state_vec` holds the gridded potential
    vorticity `pv[level, x, y]`; the generator itself steps in spectral space (dabench/data/_sqgturb.py:441-504), so
    the wrapper transforms in (`fft2_2dto1d`, :420-426) and `generate` transforms out.  The returned trajectory keeps
    the reference's scan quirk: its first row is one step past `state_vec`."""

    def __init__(self, system_dim=None, delta_t=None, model_obj=None):
        if model_obj is None:
            from ..data import SQGTurb
            model_obj = SQGTurb(delta_t=delta_t or 900)
        super().__init__(system_dim=system_dim or model_obj.system_dim,
                         delta_t=delta_t or model_obj.delta_t, model_obj=model_obj)

    def forecast(self, state_vec, n_steps=2):
        x0 = self.model_obj.fft2_2dto1d(_xr.var_array(state_vec, 'pv'))
        new_vec = self.model_obj.generate(x0=x0, n_steps=n_steps)
        return new_vec.isel(time=-1), new_vec
