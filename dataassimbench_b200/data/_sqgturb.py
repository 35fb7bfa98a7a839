"""Surface quasi-geostrophic turbulence generator with the reference's constructor / generate() contract.

Mirrors `dabench.data.SQGTurb` (dabench/data/_sqgturb.py:52-548): same arguments and defaults, the same constant
tables (built in the reference's dtypes: float32 for precision='single', wavenumbers cast through complex64 at
either precision, :215-230), the same state conventions (x0 = rfft2(pv).ravel(), complex; `values` in grid
space) and the same quirks:
  * `integrate` returns the state AFTER every one of its n_steps + 1 scan passes, so `generate(n_steps)` starts one
    step past x0 and ends n_steps + 1 steps past it (:486-496);
  * the equilibrium jet varies along axis 1 (:198-199) and only `symmetric=True` is constructible (:383-396 indexes
    a 1-D array with [1, :]);
  * `_spectrunc` does not undo `_specpad`'s 2.25 (:293-325);
  * `dealias=False` never defines the tendency (:521-533) and raises.
The time stepping runs on the GPU for all members at once (`dab_sqg_forecast_c128`, csrc/sqg_forecast.cu) in FP64
whatever `precision` says: 'single' only decides how the tables are rounded, as it does for the reference's
constants; the reference's float32 trajectory then differs from this one at the 1e-6 level.
"""
import os

import numpy as np

from .. import _xr
from ._lorenz96 import Data

_DEFAULT_PV = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), '_suppl_data',
                           'sqgturb_57600steps.npy')


class SQGTurb(Data):
    def __init__(self, pv=None, f=1.0e-4, nsq=1.0e-4, L=20.0e6, H=10.0e3, U=30.0, r=0.0, tdiab=10.0 * 86400,
                 diff_order=8, diff_efold=86400. / 3, symmetric=True, dealias=True, precision='single', tstart=0,
                 delta_t=900, store_as_jax=False, **kwargs):
        self.is_spectral = True
        super().__init__(delta_t=delta_t, store_as_jax=store_as_jax, **kwargs)
        self.coord_names = ['level', 'x', 'y']
        self.var_names = ['pv']
        if pv is None:
            pv = np.load(_DEFAULT_PV)
        pv = np.asarray(pv)
        self.original_dim = pv.shape
        if pv.shape[0] != 2:
            raise ValueError('1st dim of pv should be 2')
        N = pv.shape[1]
        if N % 2:
            raise ValueError('2nd dim of pv (N) must be even(powers of 2 are fastest)')
        if N % 4:
            raise ValueError('the dealiasing grid 3N/2 must be even: N = {} is not a multiple of 4'.format(N))
        self.N = N
        self.system_dim = 2 * N ** 2
        if precision == 'single':
            ft = np.float32
        elif precision == 'double':
            ft = np.float64
        else:
            raise ValueError('Precision must be "single" or "double"')
        if delta_t is None:
            raise ValueError('must specify time step delta_t = {}'.format(delta_t))
        if diff_efold is None:
            raise ValueError('must specify efolding time scale for diffusion')
        self._ft = ft
        self.pvspec = self.fft2(pv)
        self.x0 = self.pvspec.ravel()
        self.Nv, self.Nx, self.Ny = self.pvspec.shape
        self._x0_gridded = None
        self.nsq, self.f, self.H, self.U, self.L = ft(nsq), ft(f), ft(H), ft(U), ft(L)
        self.delta_t = ft(delta_t)
        self.dealias = dealias
        self.ekman = not (r < 1.0e-10)
        self.r, self.tdiab = ft(r), ft(tdiab)
        self.t = tstart
        self.symmetric = symmetric
        # equilibrium jet the thermal relaxation pulls towards
        y = np.arange(0, self.L, self.L / ft(N), dtype=ft)
        pi = ft(np.pi)
        l = ft(2.0) * pi / self.L
        mu = ft(l * ft(np.sqrt(nsq)) * self.H / self.f)
        if symmetric:
            jet = self._symmetric_pvbar(mu, self.U, l, self.H, y)
        else:
            jet = self._asymmetric_pvbar(mu, self.U, l, self.H, y)
        self.pvbar = np.expand_dims(jet.astype(ft), 1) * np.ones((2, N, N), ft)
        self.pvspec_eq = self.fft2(self.pvbar)
        # wavenumbers of the N grid and of the 3N/2 dealiasing grid
        self.ksqlsq, self.ik, self.il = self._wavenumbers(N, N // 2 + 1, pi)
        if dealias:
            _, self.ik_pad, self.il_pad = self._wavenumbers(3 * N // 2, 3 * N // 4 + 1, pi)
        mu = np.sqrt(self.ksqlsq) * np.sqrt(self.nsq) * self.H / self.f
        mu = mu.clip(np.finfo(mu.dtype).eps)
        self.Hovermu = (self.H / mu).astype(ft)
        with np.errstate(over='ignore'):
            self.tanhmu = np.tanh(mu.astype(np.float64)).astype(ft)
            self.sinhmu = np.sinh(mu.astype(np.float64)).astype(ft)
        self.diff_order, self.diff_efold = ft(diff_order), ft(diff_efold)
        ktot = np.sqrt(self.ksqlsq)
        ktotcutoff = ft(pi * ft(N) / self.L)
        self.hyperdiff = np.exp((-self.delta_t / self.diff_efold) * (ktot / ktotcutoff) ** self.diff_order).astype(ft)
        self._device_token = object()

    def _wavenumbers(self, n, n_half, pi):
        ft = self._ft
        k, l = np.meshgrid((n * np.fft.fftfreq(n))[0:n_half], n * np.fft.fftfreq(n))
        k = (ft(2.0) * pi * k.astype(ft) / self.L).astype(ft)
        l = (ft(2.0) * pi * l.astype(ft) / self.L).astype(ft)
        return k ** 2 + l ** 2, (1.0j * k).astype(np.complex64), (1.0j * l).astype(np.complex64)

    @staticmethod
    def _symmetric_pvbar(mu, U, l, H, y):
        half = mu.dtype.type(0.5)
        return -(mu * half * U / (l * H)) * np.cosh(half * mu) * np.cos(l * y) / np.sinh(half * mu)

    @staticmethod
    def _asymmetric_pvbar(mu, U, l, H, y):
        pvbar = -(mu * U / (l * H)) * np.cos(l * y) / np.sinh(mu)
        pvbar[1, :] = pvbar[0, :] * np.cosh(mu)          # IndexError on the 1-D jet, as in the reference
        return pvbar

    # ---- grid <-> spectral helpers (same names as the reference, :398-438)
    def fft2(self, pv):
        pv = np.asarray(pv)
        out = np.fft.rfft2(pv)
        return out.astype(np.complex64) if pv.dtype == np.float32 else out

    def ifft2(self, pvspec):
        pvspec = np.asarray(pvspec)
        out = np.fft.irfft2(pvspec)
        return out.astype(np.float32) if pvspec.dtype == np.complex64 else out

    def map2dto1d(self, pv):
        return np.asarray(pv).ravel()

    def map1dto2d(self, x):
        return np.reshape(x, (self.Nv, self.Nx, self.Ny))

    def fft2_2dto1d(self, pv):
        return self.map2dto1d(self.fft2(pv))

    def ifft2_2dto1d(self, pvspec):
        return self.map2dto1d(self.ifft2(pvspec))

    def map1dto2d_fft2(self, x):
        return self.fft2(self.map1dto2d(x))

    def map1dto2d_ifft2(self, x):
        return self.ifft2(self.map1dto2d(x))

    @property
    def x0_gridded(self):
        return None if self.x0 is None else self.map1dto2d_ifft2(self.x0)

    # ---- device side
    def _setup_device(self):
        from .._runtime import handle
        if not self.dealias:
            raise UnboundLocalError("cannot access local variable 'dpvspecdt': the reference's rhs only defines "
                                    "the tendency when dealias=True")
        H = handle()
        N = self.N
        # the handle keeps ONE generator's tables and DFT matrices: upload them only when another generator (or
        # nobody) set it up last -- a cycler calls this once per cycle
        if getattr(H, '_sqg_owner', None) is self._device_token:
            return H
        H.sqg_setup(N, float(self.delta_t), float(self._ft(1.0) / self.tdiab), float(self.r), int(self.ekman),
                    int(bool(self.symmetric)),
                    kx=self.ik_pad[0, :N // 2 + 1].imag, ly=self.il[:, 0].imag, hovermu=self.Hovermu,
                    tanhmu=self.tanhmu, sinhmu=self.sinhmu, ksqlsq=self.ksqlsq, hyperdiff=self.hyperdiff,
                    pvspec_eq=self.pvspec_eq)
        H._sqg_owner = self._device_token
        return H

    def forecast_members(self, spec, n_steps, return_grid_traj=False):
        """`n_steps` RK4 steps of every member at once.  spec: [k, 2, N, N/2+1] complex (NumPy or CUDA tensor).
        Returns the final spectral states (same kind as the input) and, if asked, the grid-space trajectory
        [n_steps, k, 2, N, N] of the states after step 1..n_steps (CUDA tensor)."""
        import torch
        H = self._setup_device()
        on_dev = isinstance(spec, torch.Tensor)
        x = spec if on_dev else torch.from_numpy(np.ascontiguousarray(np.asarray(spec, dtype=np.complex128)))
        x = x.to(device='cuda', dtype=torch.complex128).contiguous()
        if x.ndim != 4 or tuple(x.shape[1:]) != (2, self.N, self.N // 2 + 1):
            raise ValueError('spectral members must have shape (k, 2, %d, %d), got %s'
                             % (self.N, self.N // 2 + 1, tuple(x.shape)))
        k = x.shape[0]
        out = torch.empty_like(x)
        traj = torch.empty((max(n_steps, 1), k, 2, self.N, self.N), dtype=torch.float64, device='cuda') \
            if return_grid_traj else None
        H.sqg_forecast(x, out, traj, k, n_steps, torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().synchronize()
        res = out if on_dev else out.cpu().numpy()
        return (res, traj[:n_steps]) if return_grid_traj else res

    def forecast_members_grid(self, grid, n_steps, return_grid_traj=False):
        """The same for GRIDDED members [k, 2, N, N] that live on the device (CUDA tensor in, CUDA tensors out):
        rfft2 in, `n_steps` RK4 steps, irfft2 out, all in `dab_sqg_forecast_grid_f64`."""
        import torch
        H = self._setup_device()
        x = grid.to(device='cuda', dtype=torch.float64).contiguous()
        if x.ndim != 4 or tuple(x.shape[1:]) != (2, self.N, self.N):
            raise ValueError('gridded members must have shape (k, 2, %d, %d), got %s' % (self.N, self.N, tuple(x.shape)))
        k = x.shape[0]
        out = torch.empty_like(x)
        traj = torch.empty((max(n_steps, 1), k, 2, self.N, self.N), dtype=torch.float64, device='cuda') \
            if return_grid_traj else None
        H.sqg_forecast_grid(x, out, None, traj, k, n_steps, torch.cuda.current_stream().cuda_stream)
        return (out, traj[:n_steps]) if return_grid_traj else out

    def integrate(self, f, x0, t_final, delta_t=None, include_x0=True, t=None, **kwargs):
        """:441-504.  `f` is unused, as in the reference."""
        pvspec = self.map1dto2d(x0)
        n_steps = int(t_final / self.delta_t)
        if not n_steps * delta_t == t_final:
            raise ValueError('Cannot have remainder in nsteps = {}, delta_t = {}, t_final = {}, and n_steps * '
                             'delta_t = {}'.format(n_steps, delta_t, t_final, n_steps * delta_t))
        if delta_t is None:
            delta_t = self.delta_t
        if t is None:
            t = self.t
        if include_x0:
            n_steps = n_steps + 1
        times = t + np.arange(n_steps) * delta_t
        spec, traj = self.forecast_members(np.asarray(pvspec)[None], n_steps, return_grid_traj=True)
        values = traj[:, 0].cpu().numpy()
        self.pvspec = spec[0]
        self.t = times[-1]
        return values, times

    def rhs(self, pvspec):
        """Host-side tendency (API compatibility and spot checks; the kernels never call it), :506-548."""
        pvspec = np.asarray(pvspec)
        N, h = self.N, self.N // 2
        psispec = np.array([self.Hovermu * ((pvspec[1] / self.sinhmu) - (pvspec[0] / self.tanhmu)),
                            self.Hovermu * ((pvspec[1] / self.tanhmu) - (pvspec[0] / self.sinhmu))],
                           dtype=pvspec.dtype)

        def deriv(spec):
            pad = np.zeros((2, 3 * N // 2, 3 * N // 4 + 1), dtype=spec.dtype)
            pad[:, :h, :h] = 2.25 * spec[:, :h, :h]
            pad[:, -h:, :h] = 2.25 * spec[:, -h:, :h]
            pad[:, :h, h] = np.conjugate(2.25 * spec[:, :h, -1])
            pad[:, -h:, h] = np.conjugate(2.25 * spec[:, -h:, -1])
            return self.ifft2(self.ik_pad * pad), self.ifft2(self.il_pad * pad)

        if not self.dealias:
            raise UnboundLocalError("cannot access local variable 'dpvspecdt'")
        psix, psiy = deriv(psispec)
        pvx, pvy = deriv(pvspec)
        full = self.fft2(psix * pvy - psiy * pvx)
        jac = np.zeros((2, N, h + 1), dtype=full.dtype)
        jac[:, :h, :h] = full[:, :h, :h]
        jac[:, -h:, :h] = full[:, -h:, :h]
        d = (self._ft(1.0) / self.tdiab) * (self.pvspec_eq - pvspec) - jac
        if self.ekman:
            d[0] = d[0] + self.r * self.ksqlsq * psispec[0]
            if self.symmetric:
                d[1] = d[1] + (-1 * self.r * self.ksqlsq * psispec[1])
        self.u, self.v = -psiy, psix
        return d

    def generate(self, n_steps=None, t_final=None, x0=None, **kwargs):
        """Data.generate for a generator with its own integrate (dabench/data/_data.py:121-189)."""
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
        y, t = self.integrate(self.rhs, x0, t_final, self.delta_t, **kwargs)
        y_out = np.asarray(y).reshape((t.shape[0],) + tuple(self.original_dim))
        coords = {'time': t}
        for name, dim in zip(self.coord_names, self.original_dim):
            coords[name] = np.arange(dim)
        return _xr.new_dataset({'pv': (('time',) + tuple(self.coord_names), y_out)}, coords=coords,
                               attrs={'store_as_jax': self.store_as_jax, 'system_dim': self.system_dim,
                                      'delta_t': self.delta_t})
