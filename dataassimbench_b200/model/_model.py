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
    """Wraps a `data.SQGTurb` generator behind the same plugin point.  `state_vec` holds the gridded potential
    vorticity -- `pv[level, x, y]`, or the cyclers' flattened `x[index]` (dabench/dacycler/_etkf.py:213 knows no other
    name) -- while the generator steps in spectral space (dabench/data/_sqgturb.py:441-504), so the wrapper
    transforms in (`fft2`, :398-404) and out (`ifft2`, :496).  The trajectory follows the CYCLERS' contract, the one
    `Lorenz96Model` has: `n_steps` points at spacing delta_t, the first being `state_vec` itself -- NOT the scan quirk
    of `SQGTurb.generate()` (whose rows start one step late and end one step late): a cycler that forecast
    `steps_per_window + 1` steps per window of `steps_per_window - 1` would run ahead of the observations."""

    def __init__(self, system_dim=None, delta_t=None, model_obj=None):
        if model_obj is None:
            from ..data import SQGTurb
            model_obj = SQGTurb(delta_t=delta_t or 900)
        super().__init__(system_dim=system_dim or model_obj.system_dim,
                         delta_t=delta_t or model_obj.delta_t, model_obj=model_obj)

    def forecast(self, state_vec, n_steps=2):
        import numpy as np
        g = self.model_obj
        gridded = 'pv' in state_vec.data_vars
        pv = np.asarray(_xr.var_array(state_vec, 'pv' if gridded else 'x'), dtype=np.float64).reshape(g.original_dim)
        n = int(n_steps) - 1
        rows = pv[None]
        if n > 0:
            _, traj = g.forecast_members(np.fft.rfft2(pv)[None], n, return_grid_traj=True)
            rows = np.concatenate([rows, traj[:, 0].cpu().numpy()], axis=0)
        times = np.arange(rows.shape[0]) * float(g.delta_t)
        if gridded:
            coords = {'time': times}
            coords.update({name: np.arange(d) for name, d in zip(g.coord_names, g.original_dim)})
            ds = _xr.new_dataset({'pv': (('time',) + tuple(g.coord_names), rows)}, coords=coords)
        else:
            ds = _xr.new_dataset({'x': (('time', 'index'), rows.reshape(rows.shape[0], -1))}, coords={'time': times})
        return ds.isel(time=-1), ds

    def _forecast_members_device(self, X, n_steps, want_rows=False):
        """All members at once on the device, for the cyclers: X [k, 2 N N] flattened gridded states (CUDA) ->
        (last row, first row, every row [n_steps, k, 2 N N] or None) of the trajectories `forecast` returns member by
        member: rfft2, n_steps - 1 RK4 steps, irfft2 in one call (`dab_sqg_forecast_grid_f64`)."""
        import torch
        g = self.model_obj
        k = X.shape[0]
        n = int(n_steps) - 1
        res = g.forecast_members_grid(X.reshape(k, 2, g.N, g.N), n, return_grid_traj=want_rows)
        last, traj = res if want_rows else (res, None)
        rows = torch.cat([X[None], traj.reshape(traj.shape[0], k, -1)], dim=0) if want_rows else None
        return last.reshape(k, -1), X, rows
