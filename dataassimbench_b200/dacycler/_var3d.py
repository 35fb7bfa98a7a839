"""3D-Var cycler with the reference's constructor and `cycle()` signature.

Mirrors `dabench.dacycler.Var3D` (dabench/dacycler/_var3d.py:18-112) under `DACycler.cycle`
(dabench/dacycler/_dacycler.py:186-290).  Accelerated contract: a Model wrapping one of the package's generators (`Lorenz96Model`: the RK4 forecast kernel
with one member; `Lorenz63Model` / `SQGTurbModel`: their device forecast through `_forecast_members_device`), default H
(index gather), default R (`obs_error_sd**2 * I`).  With the default B (identity,
`DACycler._calc_default_B`) the system `I + B H^T R^-1 H` the reference solves by conjugate
gradients (tol 1e-5) is diagonal, and `dab_var3d_analysis_f64` solves it exactly; with a
user-supplied dense B, `dab_var3d_analysis_b_f64` runs the reference's conjugate gradients on the
device (same start, same stopping rule; one pass over B per iteration).  The forecast is the
Lorenz-96 RK4 kernel with one member.  User-supplied H / R or another model raise
`NotImplementedError` -- there is no CPU fallback.
"""
import numpy as np

from .. import _xr
from ._dacycler import DACycler


class Var3D(DACycler):
    _in_4d = False
    _uses_ensemble = False
    _device = 'cuda'

    def __init__(self, system_dim, delta_t, model_obj, B=None, R=None, H=None, h=None):
        super().__init__(system_dim=system_dim, delta_t=delta_t, model_obj=model_obj, B=B, R=R, H=H, h=h)

    def _handle(self):
        from .._runtime import handle
        return handle()

    # ------------------------------------------------------------------ marshalling (CPU-testable)
    def _plan(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd=None,
              analysis_window=0.2, analysis_time_in_window=None):
        if self.h is not None and self.H is None:
            # dabench/dacycler/_dacycler.py:103-104
            raise ValueError('Only linear obs operators (H) are supported right now.')
        if self.H is not None or self.R is not None:
            raise NotImplementedError('user-supplied H / R matrices are outside the accelerated path '
                                      '(default index-gather H and obs_error_sd**2 * I only)')
        gen = self._lorenz96_generator()
        if gen is None and not callable(getattr(self.model_obj, '_forecast_members_device', None)):
            raise NotImplementedError('the accelerated Var3D.cycle() needs model_obj to be a Model wrapping a '
                                      'dataassimbench_b200.data generator (model.Lorenz96Model, Lorenz63Model, '
                                      'SQGTurbModel), i.e. one with a device forecast')
        w = self._prepare_cycle(input_state, start_time, obs_vector, n_cycles, obs_error_sd,
                                analysis_window, analysis_time_in_window)
        if self._data_vars != ['x'] or self._observed_vars != ['x']:
            raise NotImplementedError("single-variable states named 'x' only")
        if np.ndim(self.obs_error_sd) != 0:
            raise NotImplementedError('scalar obs_error_sd only')
        x0 = np.array(_xr.var_array(input_state, 'x'), dtype=np.float64, order='C', copy=True)   # never the caller's array
        if x0.ndim != 1:
            raise ValueError("input_state['x'] must have one dimension (the state)")
        vals = np.asarray(_xr.var_array(obs_vector, 'x'), dtype=np.float64)
        vals = np.ascontiguousarray(vals.reshape(len(w['obs_times']), -1))
        sysidx = np.asarray(_xr.var_array(obs_vector, 'system_index'))
        sysidx = np.ascontiguousarray(sysidx.reshape(vals.shape), dtype=np.int64)
        B = None
        if self.B is not None:
            B = np.ascontiguousarray(self.B, dtype=np.float64)
            if B.shape != (x0.size, x0.size):
                raise ValueError('B must be (system_dim, system_dim) = (%d, %d), got %s' % (x0.size, x0.size, B.shape))
        return dict(x0=x0, vals=vals, sysidx=sysidx, B=B, padded=np.asarray(w['padded'], dtype=np.int64),
                    stationary=w['stationary'], n_steps=self.steps_per_window - 1,
                    model_dt=float(gen.delta_t) if gen is not None else float(self.delta_t),
                    forcing=float(gen.forcing_term) if gen is not None else 0.0, generic=gen is None,
                    sd=float(self.obs_error_sd),
                    state_dim=_xr.var_dims(input_state, 'x')[0])

    @staticmethod
    def _rows_of_cycle(p, c):
        """(loc, val) of the valid observation rows of cycle c: slots with padded index > 0
        (`obs_time_mask`, dabench/dacycler/_dacycler.py:119) and, for moving observers, non-NaN values
        (`_obs_loc_masks`, :262-268).  Zeroing the other rows of H (_var3d.py:88-89) = dropping them."""
        idx = p['padded'][c] if p['padded'].ndim == 2 and p['padded'].shape[1] else np.zeros(0, dtype=np.int64)
        t = idx[idx > 0] - 1
        if len(t) == 0 or p['vals'].shape[1] == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        loc, val = p['sysidx'][t].ravel(), p['vals'][t].ravel()
        if not p['stationary']:
            keep = ~np.isnan(val)
            loc, val = loc[keep], val[keep]
        return np.ascontiguousarray(loc), np.ascontiguousarray(val)

    # ------------------------------------------------------------------ the call
    def cycle(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd=None,
              analysis_window=0.2, analysis_time_in_window=None, return_forecast=False):
        """Analysis + forecast, `n_cycles` times.  Returns `x[cycle, index]` (analyses) or, with
        return_forecast=True, `x[cycle, cycle_timestep, index]` with the analysis at cycle_timestep 0 and
        the last forecast point (the next background) dropped (dabench/dacycler/_dacycler.py:281-290)."""
        import torch
        p = self._plan(input_state, start_time, obs_vector, n_cycles, obs_error_sd,
                       analysis_window, analysis_time_in_window)
        H = self._handle()
        dev = self._device
        N, S = p['x0'].size, p['n_steps']
        stream = torch.cuda.current_stream().cuda_stream if dev == 'cuda' else None
        xb = torch.from_numpy(p['x0']).to(dev)
        xa = torch.empty_like(xb)
        traj = torch.empty((max(S, 1), 1, N), dtype=torch.float64, device=dev)
        full = np.empty((n_cycles, max(S, 1), N))
        Bd = None if p['B'] is None else torch.from_numpy(p['B']).to(dev)
        self.cg_iterations = []                                  # per cycle, when B is user-supplied
        for c in range(n_cycles):
            loc, val = self._rows_of_cycle(p, c)
            d_loc = torch.from_numpy(loc).to(dev) if len(loc) else None
            d_val = torch.from_numpy(val).to(dev) if len(loc) else None
            if Bd is None:
                H.var3d_analysis(xb, xa, N, d_loc, d_val, len(loc), p['sd'], stream)
            else:
                # dabench/dacycler/_var3d.py:105-106: cg(A, b1, x0=xb, tol=1e-05, maxiter=1000)
                self.cg_iterations.append(H.var3d_analysis_b(xb, xa, N, d_loc, d_val, len(loc), p['sd'], Bd,
                                                             1e-5, 1000, stream))
            full[c, 0] = xa.cpu().numpy()
            if p['generic']:
                # another generator behind the same plugin point: its whole-"ensemble" device forecast with one member
                # (model.Lorenz63Model / SQGTurbModel); rows = the trajectory `forecast` would return
                nxt, _, rows = self.model_obj._forecast_members_device(xa.view(1, N), self.steps_per_window, True)
                if S > 1:
                    full[c, 1:] = rows[1:S, 0].cpu().numpy()
                xb.copy_(nxt.reshape(N))
            elif S > 0:
                # members are rows: one member of N points; the next background is the last step
                H.l96_forecast(xa.view(1, N), xb.view(1, N), traj, N, 1, S, p['model_dt'], p['forcing'], stream)
                if S > 1:
                    full[c, 1:] = traj[:S - 1, 0].cpu().numpy()
            else:
                xb.copy_(xa)
        dim = p['state_dim']
        if return_forecast:
            return _xr.new_dataset({'x': (('cycle', 'cycle_timestep', dim), full)})
        return _xr.new_dataset({'x': (('cycle', dim), np.ascontiguousarray(full[:, 0]))})
