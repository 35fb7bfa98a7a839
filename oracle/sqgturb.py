"""Oracle: surface quasi-geostrophic turbulence generator (TEST INFRASTRUCTURE ONLY).

NumPy restatement of `dabench.data.SQGTurb` (dabench/data/_sqgturb.py): two-boundary constant-PV f-plane QG,
FFT spectral collocation, 2/3-rule dealiasing on a 3N/2 grid, classical RK4 with the hyperdiffusion applied as an
integrating factor after the step.  Function by function:
  * tables()            -- __init__, _sqgturb.py:83-248 (wavenumbers cast through complex64 EVEN at precision
                           'double', :215-216, :229-230; mu clipped at eps, sinh/tanh in float64, :232-237;
                           hyperdiffusion factor :246-248; equilibrium jet :175-201, :366-396)
  * invert()            -- _invert, :258-270
  * specpad/spectrunc   -- :293-325 (2.25 = (3/2)^2 on the way in, NO 1/2.25 on the way out: the reference's
                           Jacobian is 2.25 times the collocation value; kept)
  * rhs()               -- :506-548 (dealiased branch; the reference's non-dealiased branch never defines the
                           tendency, so dealias=False raises there and is not restated)
  * rk4_step()          -- _rk4, :345-364
  * integrate()         -- :441-504: `lax.scan(self._rk4, pvspec, length=n_steps + 1)` returns the state AFTER
                           every step, so `values[0]` is one step past x0 (not x0) and the last of the
                           n_steps + 1 rows is n_steps + 1 steps past it; times = t + arange(n_steps + 1) dt.  Kept.
FFT conventions are numpy.fft's (= jax.numpy.fft's): unnormalised forward, 1/n inverse, c2r ignores the imaginary
parts of the zero and Nyquist columns.

Pinned by the one numerical golden the reference holds for this generator, tests/observer_base_test.py:240-261
(default SQGTurb(), 10 steps, seeded Observer: observed value 2128.20749725, system_index 10130, error draw
25.714826330507496, times [2700, 9000]) -- tests/test_oracle_golden.py.
"""
import numpy as np


def tables(N, precision='single', f=1.0e-4, nsq=1.0e-4, L=20.0e6, H=10.0e3, U=30.0, r=0.0, tdiab=10.0 * 86400,
           diff_order=8, diff_efold=86400. / 3, symmetric=True, delta_t=900):
    """Every constant array of the reference's __init__, in the dtypes it ends up with there."""
    if N % 2:
        raise ValueError('2nd dim of pv (N) must be even(powers of 2 are fastest)')
    if precision == 'single':
        dtype = np.float32
    elif precision == 'double':
        dtype = np.float64
    else:
        raise ValueError('Precision must be "single" or "double"')
    T = dict(N=N, dtype=dtype, symmetric=symmetric, ekman=not (r < 1.0e-10))
    T['nsq'] = dtype(nsq); T['f'] = dtype(f); T['H'] = dtype(H); T['U'] = dtype(U); T['L'] = dtype(L)
    T['delta_t'] = dtype(delta_t); T['r'] = dtype(r); T['tdiab'] = dtype(tdiab)
    y = np.arange(0, T['L'], T['L'] / dtype(N), dtype=dtype)
    pi = dtype(np.pi)
    l = dtype(2.0) * pi / T['L']
    # :179 mixes the python float nsq with the typed H and f (weak-typed in jax: the result has `dtype`)
    mu = dtype(l * dtype(np.sqrt(nsq)) * T['H'] / T['f'])
    if not symmetric:
        # _asymmetric_pvbar (:383-396) indexes its 1-D result with [1, :]: the reference raises here
        raise IndexError('Too many indices: the reference\'s asymmetric jet is not constructible')
    pvbar = (-(mu * dtype(0.5) * T['U'] / (l * T['H'])) * np.cosh(dtype(0.5) * mu) * np.cos(l * y)
             / np.sinh(dtype(0.5) * mu)).astype(dtype)
    # :198-199: expand_dims(pvbar, 1) of the 1-D jet is (N, 1); times ones((2, N, N)) it varies along axis 1
    pvbar = np.expand_dims(pvbar, 1) * np.ones((2, N, N), dtype)
    T['pvbar'] = pvbar
    T['pvspec_eq'] = np.fft.rfft2(pvbar)
    k = (N * np.fft.fftfreq(N))[0:(N // 2) + 1]
    lw = N * np.fft.fftfreq(N)
    k, lw = np.meshgrid(k, lw)
    k = (dtype(2.0) * pi * k.astype(dtype) / T['L']).astype(dtype)
    lw = (dtype(2.0) * pi * lw.astype(dtype) / T['L']).astype(dtype)
    T['ksqlsq'] = k ** 2 + lw ** 2
    T['ik'] = (1.0j * k).astype(np.complex64)
    T['il'] = (1.0j * lw).astype(np.complex64)
    M = 3 * N // 2
    k_pad = (M * np.fft.fftfreq(M))[0:(3 * N // 4) + 1]
    l_pad = M * np.fft.fftfreq(M)
    k_pad, l_pad = np.meshgrid(k_pad, l_pad)
    k_pad = (dtype(2.0) * pi * k_pad.astype(dtype) / T['L']).astype(dtype)
    l_pad = (dtype(2.0) * pi * l_pad.astype(dtype) / T['L']).astype(dtype)
    T['ik_pad'] = (1.0j * k_pad).astype(np.complex64)
    T['il_pad'] = (1.0j * l_pad).astype(np.complex64)
    mu2 = np.sqrt(T['ksqlsq']) * np.sqrt(T['nsq']) * T['H'] / T['f']
    mu2 = mu2.clip(np.finfo(mu2.dtype).eps)
    T['Hovermu'] = (T['H'] / mu2).astype(dtype)
    mu64 = mu2.astype(np.float64)
    with np.errstate(over='ignore'):
        T['tanhmu'] = np.tanh(mu64).astype(dtype)
        T['sinhmu'] = np.sinh(mu64).astype(dtype)
    ktot = np.sqrt(T['ksqlsq'])
    ktotcutoff = dtype(pi * dtype(N) / T['L'])
    T['hyperdiff'] = np.exp((-T['delta_t'] / dtype(diff_efold)) * (ktot / ktotcutoff) ** dtype(diff_order)).astype(dtype)
    return T


def invert(T, pvspec):
    return np.array([T['Hovermu'] * ((pvspec[1] / T['sinhmu']) - (pvspec[0] / T['tanhmu'])),
                     T['Hovermu'] * ((pvspec[1] / T['tanhmu']) - (pvspec[0] / T['sinhmu']))], dtype=pvspec.dtype)


def specpad(T, specarr):
    N = T['N']
    out = np.zeros((2, 3 * N // 2, 3 * N // 4 + 1), dtype=specarr.dtype)
    out[:, :N // 2, :N // 2] = 2.25 * specarr[:, :N // 2, :N // 2]
    out[:, -N // 2:, :N // 2] = 2.25 * specarr[:, -N // 2:, :N // 2]
    out[:, :N // 2, N // 2] = np.conjugate(2.25 * specarr[:, :N // 2, -1])
    out[:, -N // 2:, N // 2] = np.conjugate(2.25 * specarr[:, -N // 2:, -1])
    return out


def spectrunc(T, specarr):
    N = T['N']
    out = np.zeros((2, N, N // 2 + 1), dtype=specarr.dtype)
    out[:, :N // 2, :N // 2] = specarr[:, :N // 2, :N // 2]
    out[:, -N // 2:, :N // 2] = specarr[:, -N // 2:, :N // 2]
    return out


def _ifft2(T, spec):
    out = np.fft.irfft2(spec)
    # numpy promotes complex64 input to float64 output only through its own dtype rules; keep the state's class
    return out.astype(np.float32) if spec.dtype == np.complex64 else out


def _fft2(T, grid):
    out = np.fft.rfft2(grid)
    return out.astype(np.complex64) if grid.dtype == np.float32 else out


def rhs(T, pvspec):
    psispec = invert(T, pvspec)
    pad_psi, pad_pv = specpad(T, psispec), specpad(T, pvspec)
    psix, psiy = _ifft2(T, T['ik_pad'] * pad_psi), _ifft2(T, T['il_pad'] * pad_psi)
    pvx, pvy = _ifft2(T, T['ik_pad'] * pad_pv), _ifft2(T, T['il_pad'] * pad_pv)
    jacobian = psix * pvy - psiy * pvx
    jacobianspec = spectrunc(T, _fft2(T, jacobian))
    d = (T['dtype'](1.0) / T['tdiab']) * (T['pvspec_eq'] - pvspec) - jacobianspec
    if T['ekman']:
        d[0] = d[0] + T['r'] * T['ksqlsq'] * psispec[0]
        if T['symmetric']:
            d[1] = d[1] + (-1 * T['r'] * T['ksqlsq'] * psispec[1])
    return d


def rk4_step(T, x):
    dt = T['delta_t']
    k1 = dt * rhs(T, x)
    k2 = dt * rhs(T, x + 0.5 * k1)
    k3 = dt * rhs(T, x + 0.5 * k2)
    k4 = dt * rhs(T, x + k3)
    y = x + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
    return T['hyperdiff'] * y


def integrate(T, pvspec, n_scan, t0=0.0):
    """`n_scan` passes of _rk4; returns (final spectral state, grid-space values [n_scan, 2, N, N], times)."""
    vals = []
    for _ in range(n_scan):
        pvspec = rk4_step(T, pvspec)
        vals.append(pvspec)
    values = np.stack([_ifft2(T, v) for v in vals]) if vals else np.zeros((0, 2, T['N'], T['N']))
    return pvspec, values, t0 + np.arange(n_scan) * float(T['delta_t'])


def generate(pv, n_steps, precision='single', **params):
    """SQGTurb(pv=pv, precision=...).generate(n_steps=n_steps): values [n_steps + 1, 2, N, N] and times."""
    pv = np.asarray(pv)
    T = tables(pv.shape[1], precision=precision, **params)
    x0 = _fft2(T, pv)
    _, values, times = integrate(T, x0, n_steps + 1)
    return values, times
