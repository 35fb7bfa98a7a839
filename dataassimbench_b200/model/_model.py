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

    def _forecast_members_device(self, X, n_steps, want_rows=False):
        """All members at once on the device, for the cyclers: X [k, 3] (CUDA) -> (last point, first point, every
        point [n_steps, k, 3] or None) of the `n_steps`-point trajectories `forecast` returns member by member."""
        import torch
        from .._runtime import handle
        g = self.model_obj
        out = torch.empty_like(X)
        n = int(n_steps) - 1
        traj = torch.empty((max(n, 1), X.shape[0], 3), dtype=torch.float64, device=X.device) if want_rows else None
        handle().l63_forecast(X, out, traj if n > 0 else None, X.shape[0], n, float(g.delta_t), float(g.sigma),
                              float(g.rho), float(g.beta), torch.cuda.current_stream().cuda_stream)
        rows = torch.cat([X[None], traj[:n]], dim=0) if want_rows else None
        return out, X, rows


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
        if 'pv' in state_vec:
            pv = _xr.var_array(state_vec, 'pv')
        else:       # the cyclers' flattened state `x[index]` (dabench/dacycler/_etkf.py:213 knows no other name)
            pv = _xr.var_array(state_vec, 'x').reshape(self.model_obj.original_dim)
        new_vec = self.model_obj.generate(x0=self.model_obj.fft2_2dto1d(pv), n_steps=n_steps)
        if 'pv' in state_vec:
            return new_vec.isel(time=-1), new_vec
        flat = _xr.var_array(new_vec, 'pv').reshape(new_vec['pv'].shape[0], -1)
        flat_ds = _xr.new_dataset({'x': (('time', 'index'), flat)}, coords={'time': _xr.var_array(new_vec, 'time')})
        return flat_ds.isel(time=-1), flat_ds

    def _forecast_members_device(self, X, n_steps, want_rows=False):
        """All members at once on the device, for the cyclers: X [k, 2 N N] flattened gridded states (CUDA) ->
        (last row, first row, every row or None) of the trajectories `forecast` returns member by member.
        `generate(n_steps)` emits the states after steps 1 .. n_steps + 1 (the reference's scan quirk), so the first
        row is one step past X and there are n_steps + 1 rows."""
        g = self.model_obj
        k = X.shape[0]
        last, traj = g.forecast_members_grid(X.reshape(k, 2, g.N, g.N), int(n_steps) + 1, return_grid_traj=True)
        rows = traj.reshape(traj.shape[0], k, -1)
        return last.reshape(k, -1), rows[0], (rows if want_rows else None)
