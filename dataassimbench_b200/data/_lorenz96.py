"""Lorenz-96 generator with the reference's constructor / generate() contract.

Mirrors `dabench.data.Lorenz96` (dabench/data/_lorenz96.py:42-151) and the part of
`dabench.data.Data.generate` the ETKF path uses (dabench/data/_data.py:86-189): same
arguments, same default initial states, same output Dataset `x[time, index]` with attrs
`system_dim`, `delta_t`.  The integration runs on the GPU (K1, fixed-step RK4 at `delta_t`)
instead of the reference's adaptive `odeint` -- BASELINE.json's north star.
"""
import numpy as np

from .. import _xr

# spun-up defaults of the reference, dabench/data/_lorenz96.py:66-96 (data, not code)
_X0 = {
    0.05: {
        5: [3.552747, 5.173696, 5.4161925, 5.1463985, -0.62878036],
        6: [-1.277197, 0.5191239, 2.1629903, 4.641832, 5.013882, -3.1474972],
        36: [0.90061724, 2.2108543, 3.3563306, 7.0470520, 7.3828993, -2.2906365, 1.6358340,
             4.5246205, -0.8536633, 2.2018400, 2.5094680, 5.6148005, -1.7163916, -3.5827417,
             0.22293478, 1.8138107, 3.7354333, 5.9006715, -4.6722836, 0.4664867, 0.36800075,
             7.7004447, 3.0569422, -1.7238870, -2.1296368, 1.6388168, 5.1955190, 4.7863874,
             0.8382774, -4.0938597, 0.5181451, 1.2503184, 6.0076460, 7.1161866, -3.2190716,
             -2.3532054]},
    0.01: {
        5: [1.7660924, 0.29829425, 0.3133399, 4.521974, 8.680367],
        6: [4.321441, 9.965719, 3.7177398, 1.1666082, -0.26569614, 0.3100199],
        36: [1.0433275, 4.401105, 6.32247, 4.9386244, 3.4934044, 2.3050377, -1.1888806,
             0.83992016, 2.52427, 10.146448, 1.5634892, 4.289391, 5.6141367, 5.9368305,
             -1.291068, -2.2805283, 3.365577, 6.793376, 0.41429618, 3.4392211, 5.739969,
             -3.6751282, 1.2353066, 0.04940181, 1.9196223, 6.1237254, 3.356311, -2.9513268,
             1.8416065, 3.6102607, 9.518854, 2.4801574, -0.44426048, 5.1087484, -1.305786,
             -0.79453385]},
}


class Data:
    """Minimal base: holds the attributes `Data.__init__` sets for a 1-D system."""

    def __init__(self, system_dim=3, delta_t=0.01, store_as_jax=False, x0=None, **kwargs):
        self.system_dim = system_dim
        self.original_dim = (system_dim,)
        self.delta_t = delta_t
        self.store_as_jax = store_as_jax
        self.coord_names = ['index']
        self.var_names = ['x']
        self.x0 = x0

    def generate(self, n_steps=None, t_final=None, x0=None, **kwargs):
        """Trajectory of n_steps points at spacing delta_t, initial state included
        (dabench/data/_data.py:121-125, dabench/data/_utils.py:44)."""
        if n_steps is not None:
            t_final = n_steps * self.delta_t
        elif t_final is not None:
            n_steps = int(t_final / self.delta_t)
        else:
            raise TypeError('Either n_steps or t_final must be supplied as an input argument.')
        if x0 is None:
            if self.x0 is None:
                raise TypeError('Initial condition is None, x0 = {}. it must either be provided as an '
                                'argument or set as an attribute in the model object.'.format(x0))
            x0 = self.x0
        t = np.arange(0.0, t_final - self.delta_t / 2, self.delta_t)
        y = self._integrate(np.asarray(x0, dtype=np.float64).reshape(1, -1), len(t))[0]
        return _xr.new_dataset(
            {'x': (('time', 'index'), y)},
            coords={'time': t, 'index': np.arange(self.system_dim)},
            attrs={'store_as_jax': self.store_as_jax, 'system_dim': self.system_dim,
                   'delta_t': self.delta_t})

    def _integrate(self, X0, n_points):
        raise NotImplementedError('a generator integrates on the GPU through its own forecast kernel')


class Lorenz96(Data):
    def __init__(self, forcing_term=8., delta_t=0.05, x0=None, system_dim=36, values=None,
                 store_as_jax=False, **kwargs):
        super().__init__(system_dim=system_dim, delta_t=delta_t, store_as_jax=store_as_jax, **kwargs)
        if system_dim < 4:                                   # dabench/data/_lorenz96.py:57-59
            raise ValueError('system_dim is {}, for Lorenz96 it must be >4'.format(system_dim))
        self.forcing_term = forcing_term
        if x0 is None:                                       # :98-116
            table = _X0.get(delta_t, _X0[0.01])
            if system_dim in table:
                x0 = np.array(table[system_dim], dtype=np.float64)
            else:
                x0 = np.zeros(system_dim)
                x0[0] = 0.01
        self.x0 = np.asarray(x0, dtype=np.float64)

    def rhs(self, x, t=None):
        """Host-side tendency (API compatibility; the kernels never call it)."""
        x = np.asarray(x, dtype=np.float64)
        xm2, xm1, xp1 = np.roll(x, 2, -1), np.roll(x, 1, -1), np.roll(x, -1, -1)
        return (xp1 - xm2) * xm1 - x + self.forcing_term

    def _integrate(self, X0, n_points):
        """[k, N] -> [k, n_points, N] on the GPU (K1 with trajectory output)."""
        import torch
        from .._runtime import handle
        H = handle()
        k, N = X0.shape
        x_in = torch.from_numpy(np.ascontiguousarray(X0)).cuda()
        x_out = torch.empty_like(x_in)
        n = n_points - 1
        traj = torch.empty((max(n, 1), k, N), dtype=torch.float64, device='cuda')
        H.l96_forecast(x_in, x_out, traj if n > 0 else None, N, k, n, float(self.delta_t),
                       float(self.forcing_term), torch.cuda.current_stream().cuda_stream)
        out = np.empty((k, n_points, N))
        out[:, 0] = X0
        if n > 0:
            out[:, 1:] = traj.permute(1, 0, 2).cpu().numpy()
        return out
