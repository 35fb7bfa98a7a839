"""Surface-QG ensemble forecast on the GPU (dab_sqg_forecast_c128, csrc/sqg_forecast.cu) against the NumPy oracle
(oracle/sqgturb.py).  FP64 on both sides with the same tables; the device does its pruned transforms as DFT matrix
products, the oracle as FFTs: agreement to <= 1e-10 of the field's magnitude over a few RK4 steps."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import sqgturb as o_sqg
from dataassimbench_b200.data import SQGTurb
from test_sqgturb_cpu import DEFAULT_PV, reference_observer_draws


def members(N, k, seed):
    rng = np.random.default_rng(seed)
    x, y = np.meshgrid(np.arange(N) * 2 * np.pi / N, np.arange(N) * 2 * np.pi / N)
    pv = 300.0 * rng.standard_normal((k, 2, N, N))
    pv[:, 1] += 2000.0 * (np.sin(x / 2) ** 40 * np.sin(y) ** 20)         # the blob of the reference's own test
    pv -= pv.mean(axis=(2, 3), keepdims=True)
    return pv


@pytest.mark.parametrize("N,k,n_steps,r", [(16, 1, 3, 0.0), (32, 5, 4, 0.0), (64, 3, 2, 1.0e-6), (96, 4, 3, 0.0),
                                           (40, 2, 2, 5.0e-7)])
def test_sqg_forecast_matches_oracle(N, k, n_steps, r):
    pv = members(N, k, seed=N + k)
    g = SQGTurb(pv=pv[0], precision='double', r=r)
    T = o_sqg.tables(N, 'double', r=r)
    spec0 = np.fft.rfft2(pv)                                              # [k, 2, N, N/2+1]
    spec, traj = g.forecast_members(spec0, n_steps, return_grid_traj=True)
    traj = traj.cpu().numpy()
    assert spec.shape == spec0.shape and traj.shape == (n_steps, k, 2, N, N)
    for m in range(k):
        ref_spec, ref_vals, _ = o_sqg.integrate(T, spec0[m], n_steps)
        scale = np.abs(ref_spec).max()
        assert np.abs(spec[m] - ref_spec).max() / scale < 1e-10
        assert np.abs(traj[:, m] - ref_vals).max() / np.abs(ref_vals).max() < 1e-10
    # a device-resident ensemble goes in and out without leaving the GPU, and zero steps is the identity
    dev = torch.from_numpy(spec0).cuda()
    out = g.forecast_members(dev, n_steps)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), spec)
    assert np.array_equal(g.forecast_members(spec0, 0), spec0)


def test_sqg_generate_reproduces_the_reference_observer_golden():
    """SQGTurb().generate(n_steps=10) through the package, sampled where the reference's seeded Observer samples it
    (tests/observer_base_test.py:240-261)."""
    g = SQGTurb()
    ds = g.generate(n_steps=10)
    pv = np.asarray(ds['pv'].values)
    assert pv.shape == (11, 2, 96, 96) and ds.attrs['system_dim'] == 18432
    assert np.allclose(np.asarray(ds['time'].values), np.arange(11) * 900.0)
    tsel, count, lev, xx, yy, err = reference_observer_draws(96, 11)
    assert pv[tsel[0]][lev[44], xx[44], yy[44]] + err[0, 0, 44] == pytest.approx(2128.20749725)
    ref, _ = o_sqg.generate(np.load(DEFAULT_PV), 10, precision='single')  # float32 FFT path of the reference
    assert np.abs(pv - ref).max() / np.abs(ref).max() < 2e-5
    # the generator keeps the reference's bookkeeping: final spectral state and clock
    assert g.t == 9000.0 and g.pvspec.shape == (2, 96, 49)


def test_sqg_forecast_needs_setup_and_valid_sizes():
    from dataassimbench_b200 import _native
    h = _native.Handle(0)
    try:
        with pytest.raises(_native.DabError):
            h.sqg_forecast(torch.zeros(4, device='cuda'), torch.zeros(4, device='cuda'), None, 1, 1, None)
    finally:
        h.close()
    with pytest.raises(ValueError):
        SQGTurb(pv=np.zeros((2, 6, 6)))           # 3N/2 odd
    g = SQGTurb(pv=np.ones((2, 8, 8)), precision='double')
    with pytest.raises(ValueError, match='spectral members'):
        g.forecast_members(np.zeros((1, 2, 8, 8), dtype=np.complex128), 1)
    with pytest.raises(ValueError, match='gridded members'):
        g.forecast_members_grid(torch.zeros((1, 2, 8, 5), device='cuda'), 1)


def test_sqg_model_forecast_plugin_contract():
    """model.SQGTurbModel: gridded state in, (last state, trajectory) out -- the Model.forecast contract
    (dabench/model/_model.py:10-32) around the spectral generator."""
    import dataassimbench_b200 as dab
    from dataassimbench_b200 import _xr
    N = 32
    pv = members(N, 1, seed=5)[0]
    gen = dab.data.SQGTurb(pv=pv, precision='double')
    model = dab.model.SQGTurbModel(model_obj=gen)
    assert model.system_dim == 2 * N * N and model.delta_t == 900
    state = _xr.new_dataset({'pv': (('level', 'x', 'y'), pv)},
                            coords={'level': np.arange(2), 'x': np.arange(N), 'y': np.arange(N)})
    last, traj = model.forecast(state, n_steps=4)
    # the cyclers' contract: n_steps points, the first is the state itself, then one RK4 step per point
    T = o_sqg.tables(N, 'double')
    _, ref, _ = o_sqg.integrate(T, np.fft.rfft2(pv), 3)
    got = _xr.var_array(traj, 'pv')
    assert got.shape == (4, 2, N, N) and np.array_equal(got[0], pv)
    assert np.allclose(np.asarray(traj['time'].values), np.arange(4) * 900.0)
    assert np.abs(got[1:] - ref).max() / np.abs(ref).max() < 1e-10
    assert np.array_equal(_xr.var_array(last, 'pv'), got[-1])


def test_reference_obs_sqgturb_recipe_through_the_package():
    """tests/observer_base_test.py:240-261 written against this package: default SQGTurb, 10 steps on the GPU, the
    seeded Observer on the gridded trajectory -- every golden of the reference's test."""
    import dataassimbench_b200 as dab
    sqg = dab.data.SQGTurb()
    ds = sqg.generate(n_steps=10)
    obs_vec = dab.observer.Observer(ds, random_time_density=0.3, random_location_density=0.01, error_sd=25.).observe()
    assert obs_vec['pv'].shape == (2, 204)
    assert obs_vec['time'].shape[0] == 2
    sysidx = np.asarray(obs_vec['system_index'].values)
    errors = np.asarray(obs_vec['errors'].values)
    vals = np.asarray(obs_vec['pv'].values)
    assert sysidx[0, 0, 0] == 10130
    flat = np.asarray(ds['pv'].values).reshape(11, -1)
    assert vals[1, 123] == pytest.approx(flat[10, sysidx[0, 1, 123]] + errors[0, 1, 123])
    assert vals[0, 44] == pytest.approx(2128.20749725)
    assert errors[0, 1, 187] == pytest.approx(25.714826330507496)
    assert np.allclose(np.asarray(obs_vec['time'].values), np.array([2700., 9000.]))


def test_reference_sqgturb_variable_sizes_recipe():
    """tests/data_sqgturb_test.py:16-46 written against this package (NumPy noise instead of jax.random)."""
    import dataassimbench_b200 as dab
    N = 96
    rng = np.random.default_rng(42)
    pv = (100 * rng.standard_normal((2, N, N))).astype(np.float32)
    x, y = np.meshgrid(np.arange(0, 2. * np.pi, 2. * np.pi / N), np.arange(0., 2. * np.pi, 2. * np.pi / N))
    pv[1] = pv[1] + 2000. * (np.sin(x.astype(np.float32) / 2) ** 40 * np.sin(y.astype(np.float32)) ** 20)
    for k in range(2):
        pv[k] = pv[k] - pv[k].mean()
    sqgturb = dab.data.SQGTurb(pv=pv)
    n_steps = 10
    traj = sqgturb.generate(n_steps=n_steps)
    assert traj.system_dim == 18432
    assert traj.sizes['time'] == n_steps + 1
    assert traj.dab.flatten().shape == (n_steps + 1, 18432)
    assert np.all(np.isfinite(np.asarray(traj['pv'].values)))


def test_sqg_long_run_from_the_default_state_stays_on_the_oracle_trajectory():
    """100 RK4 steps (25 h of model time) from the reference's default spin-up state, double tables: the DFT-product
    transforms and the FFT oracle stay together to <= 1e-9 of the field magnitude, and the run stays finite and keeps
    the area means it started with (the (0, 0) coefficient only feels the relaxation)."""
    g = SQGTurb(precision='double')
    T = o_sqg.tables(96, 'double')
    x0 = np.asarray(g.map1dto2d(g.x0), dtype=np.complex128)
    spec, traj = g.forecast_members(x0[None], 100, return_grid_traj=True)
    ref_spec, ref_vals, _ = o_sqg.integrate(T, x0, 100)
    assert np.all(np.isfinite(spec))
    assert np.abs(spec[0] - ref_spec).max() / np.abs(ref_spec).max() < 1e-9
    got = traj[:, 0].cpu().numpy()
    assert np.abs(got - ref_vals).max() / np.abs(ref_vals).max() < 1e-9
    assert np.abs(got[-1].mean(axis=(1, 2)) - ref_vals[-1].mean(axis=(1, 2))).max() < 1e-6


def test_sqg_gridded_entry_point_is_the_spectral_one_between_two_transforms():
    """dab_sqg_forecast_grid_f64: rfft2 in, the same RK4 steps, irfft2 out; zero steps = the identity on a real
    grid (to rounding); the optional spectral output is the state the spectral entry point returns."""
    from dataassimbench_b200._runtime import handle
    N, k, n_steps = 32, 3, 2
    pv = members(N, k, seed=21)
    g = SQGTurb(pv=pv[0], precision='double')
    dev = torch.from_numpy(pv).cuda()
    same = g.forecast_members_grid(dev, 0)
    assert np.abs(same.cpu().numpy() - pv).max() / np.abs(pv).max() < 1e-14
    out, traj = g.forecast_members_grid(dev, n_steps, return_grid_traj=True)
    spec, traj_s = g.forecast_members(np.fft.rfft2(pv), n_steps, return_grid_traj=True)
    assert np.abs(out.cpu().numpy() - np.fft.irfft2(spec)).max() / np.abs(pv).max() < 1e-13
    assert np.abs((traj - traj_s).cpu().numpy()).max() / np.abs(pv).max() < 1e-13
    assert torch.equal(traj[-1], out)
    H = handle()
    spec_out = torch.empty((k, 2, N, N // 2 + 1), dtype=torch.complex128, device='cuda')
    out2 = torch.empty_like(dev)
    H.sqg_forecast_grid(dev, out2, spec_out, None, k, n_steps, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    assert np.abs(spec_out.cpu().numpy() - spec).max() / np.abs(spec).max() < 1e-13
