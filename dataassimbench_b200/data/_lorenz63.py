"""Lorenz-63 generator with the reference's constructor / generate() contract.

Mirrors `dabench.data.Lorenz63` (dabench/data/_lorenz63.py:15-118): same parameters and defaults
(sigma 10, rho 28, beta 8/3, delta_t 0.01, x0 = [-10, -15, 21.3] after the 6000-step spin-up the
reference describes), `system_dim` forced to 3 with the reference's warning, `rhs` and `Jacobian`.
`generate()` integrates on the GPU (`dab_l63_forecast_f64`, fixed-step RK4 at `delta_t`), as the
Lorenz-96 generator does.
"""
import numpy as np

from ._lorenz96 import Data


class Lorenz63(Data):
    def __init__(self, sigma=10., rho=28., beta=8. / 3., delta_t=0.01, x0=(-10.0, -15.0, 21.3),
                 system_dim=3, values=None, store_as_jax=False, **kwargs):
        if system_dim is None:
            system_dim = 3
        elif system_dim != 3:                                 # dabench/data/_lorenz63.py:49-54
            print('WARNING: input system_dim is {}, '
                  'Lorenz63 requires system_dim=3.'.format(system_dim))
            print('Assigning system_dim to 3.')
            system_dim = 3
        super().__init__(system_dim=system_dim, delta_t=delta_t, store_as_jax=store_as_jax, **kwargs)
        self.sigma = sigma
        self.rho = rho
        self.beta = beta
        self.x0 = None if x0 is None else np.asarray(x0, dtype=np.float64)

    def rhs(self, x, t=None):
        """Host-side tendency (API compatibility; the kernel never calls it), :70-90."""
        x = np.asarray(x, dtype=np.float64)
        return np.array([self.sigma * (x[1] - x[0]),
                         self.rho * x[0] - x[1] - x[0] * x[2],
                         x[0] * x[1] - self.beta * x[2]])

    def Jacobian(self, x):
        """:92-118."""
        s, b, r = self.sigma, self.beta, self.rho
        return np.array([[-s, s, 0.0],
                         [r - x[2], -1.0, -x[0]],
                         [x[1], x[0], -b]])

    def _integrate(self, X0, n_points):
        """[k, 3] -> [k, n_points, 3] on the GPU (one thread per member, trajectory output)."""
        import torch
        from .._runtime import handle
        H = handle()
        k = X0.shape[0]
        x_in = torch.from_numpy(np.ascontiguousarray(X0)).cuda()
        x_out = torch.empty_like(x_in)
        n = n_points - 1
        traj = torch.empty((max(n, 1), k, 3), dtype=torch.float64, device='cuda')
        H.l63_forecast(x_in, x_out, traj if n > 0 else None, k, n, float(self.delta_t), float(self.sigma),
                       float(self.rho), float(self.beta), torch.cuda.current_stream().cuda_stream)
        out = np.empty((k, n_points, 3))
        out[:, 0] = X0
        if n > 0:
            out[:, 1:] = traj.permute(1, 0, 2).cpu().numpy()
        return out
