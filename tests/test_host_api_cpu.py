"""Host-side mirror of the reference interface, on CPU: window tables against the oracle's literal
restatement, Observer against the reference's known answers, error behaviour, and the arrays
`ETKF.cycle()` hands to the C ABI.  No compute kernels run here."""
import numpy as np
import pytest
import scipy.integrate

import dataassimbench_b200 as dab
from dataassimbench_b200 import _xr
from dataassimbench_b200.dacycler import _windows
from oracle import windows as o_win, l96 as o_l96, l63 as o_l63, observer as o_obs, var3d as o_var3d, threefry


@pytest.mark.parametrize("start,w,n,obs_dt,n_obs_t", [
    (0.1, 0.1, 10, 0.05, 25), (0.0, 0.2, 20, 0.2, 21), (0.5, 0.2, 7, 0.01, 200), (0.0, 0.05, 40, 0.05, 41),
    (1.0, 0.3, 5, 0.07, 40)])
def test_window_tables_equal_reference_restatement(start, w, n, obs_dt, n_obs_t):
    obs_t = np.arange(n_obs_t) * obs_dt
    centres = _windows.window_centres(start, w, n)
    assert np.array_equal(centres, o_win.get_all_times(start, w, n))
    ref = o_win.pad_time_indices(o_win.get_obs_indices(centres, obs_t, w))
    got = _windows.padded_time_table(obs_t, centres, w)
    assert got.dtype == np.int64 and np.array_equal(got, ref)


def test_window_centres_keeps_reference_arange_quirk():
    with pytest.raises(ValueError):
        _windows.window_centres(0.0, 0.1, 12)


def test_window_table_without_any_observation():
    t = _windows.padded_time_table(np.array([5.0]), _windows.window_centres(0.0, 0.2, 4), 0.2)
    assert t.shape == (4, 0)


def _nature(N, dt, n):
    x0 = o_l96.default_x0(N, dt)
    t = o_l96.time_grid(n, dt)
    y = scipy.integrate.odeint(lambda x, _t: o_l96.rhs(x), x0, t)
    return _xr.new_dataset({'x': (('time', 'index'), y)}, coords={'time': t, 'index': np.arange(N)},
                           attrs={'system_dim': N, 'delta_t': dt})


def test_observer_reference_goldens():
    """/root/reference/tests/observer_base_test.py:133-180 through the API class."""
    ds = _nature(36, 0.05, 10)
    ov = dab.observer.Observer(ds, random_time_density=0.4, random_location_density=0.2,
                               error_sd=0.7).observe()
    assert ov['x'].shape == (3, 4)
    assert np.array_equal(ov['system_index'].values, np.tile([16, 9, 31, 32], 3).reshape((1, 3, 4)))
    assert np.allclose(ov['errors'].values[0, 0], [0.5913821, -0.22663247, 0.00787655, -0.29022197])
    assert np.allclose(ov['time'].values, [0.15, 0.2, 0.45])
    assert ov.error_sd == 0.7 and ov.stationary_observers is True
    mv = dab.observer.Observer(ds, random_time_density=0.4, random_location_density=0.3, error_sd=1.2,
                               stationary_observers=False).observe()
    assert mv['x'].shape == (3, 12)
    assert mv['system_index'].values[0, 2, -1] == 32
    assert mv['errors'].values[0, 1, 4] == pytest.approx(-0.7470301078422978)
    assert mv['x'].values[2, 5] == pytest.approx(0.010236507277464613)


def test_observer_matches_oracle_bit_for_bit():
    ds = _nature(36, 0.01, 40)
    for stat in (True, False):
        a = dab.observer.Observer(ds, random_time_count=6, random_location_density=0.4, error_sd=0.3,
                                  error_bias=0.1, stationary_observers=stat, random_seed=7).observe()
        b = o_obs.observe(ds['x'].values, ds['time'].values, random_time_count=6,
                          random_location_density=0.4, error_sd=0.3, error_bias=0.1,
                          stationary_observers=stat, random_seed=7)
        assert np.array_equal(a['system_index'].values[0], b['system_index'])
        assert np.array_equal(a['x'].values, b['values'], equal_nan=True)


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):                       # dabench/data/_lorenz96.py:57-59
        dab.data.Lorenz96(system_dim=3)
    with pytest.raises(TypeError):                        # dabench/data/_data.py:127-128
        dab.data.Lorenz96(system_dim=5).generate()

    class NoForecast(dab.model.Model):
        forecast = None
    with pytest.raises(ValueError):                       # dabench/model/_model.py:28-32
        NoForecast()
    assert np.allclose(dab.data.Lorenz96(system_dim=5, delta_t=0.01).x0,
                       [1.7660924, 0.29829425, 0.3133399, 4.521974, 8.680367])
    assert np.array_equal(dab.data.Lorenz96(system_dim=40).x0, np.r_[0.01, np.zeros(39)])
    l = dab.data.Lorenz96(system_dim=6)
    assert np.array_equal(l.rhs(np.arange(6.0)), o_l96.rhs(np.arange(6.0)))


def _experiment(N=8, k=4, n_cycles=5, S=10):
    ds = _nature(N, 0.01, n_cycles * S + 1)
    obs = dab.observer.Observer(ds, times=ds['time'].values[::S], random_location_count=3,
                                error_sd=0.5, random_seed=3).observe()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01))
    X0 = ds['x'].values[0] + np.random.default_rng(0).standard_normal((k, N))
    init = _xr.new_dataset({'x': (('ensemble', 'index'), X0)}, coords={'index': np.arange(N)})
    return ds, obs, model, init


def test_lorenz63_generator_surface(capsys):
    """Constructor contract of dabench.data.Lorenz63 (dabench/data/_lorenz63.py:37-68): defaults,
    system_dim forced to 3 with the reference's warning, rhs and Jacobian."""
    g = dab.data.Lorenz63(system_dimension=3)                 # the reference's tests pass this stray kwarg
    assert (g.sigma, g.rho, g.beta, g.delta_t, g.system_dim) == (10., 28., 8. / 3., 0.01, 3)
    assert np.array_equal(g.x0, [-10.0, -15.0, 21.3])
    g5 = dab.data.Lorenz63(system_dim=5)
    assert g5.system_dim == 3 and 'requires system_dim=3' in capsys.readouterr().out
    x = np.array([1.5, -2.0, 20.0])
    assert np.allclose(g.rhs(x), o_l63.rhs(x), rtol=0, atol=0)
    eps = 1e-6
    J = np.stack([(o_l63.rhs(x + eps * e) - o_l63.rhs(x - eps * e)) / (2 * eps) for e in np.eye(3)], axis=1)
    assert np.allclose(g.Jacobian(x), J, atol=1e-7)
    with pytest.raises(TypeError):
        g.generate()                                          # neither n_steps nor t_final (dabench/data/_data.py:127-128)
    m = dab.model.Lorenz63Model(model_obj=g)
    assert m.system_dim == 3 and m.delta_t == 0.01 and callable(m.forecast)


def test_cycle_plan_marshalling():
    N, k, n_cycles, S = 8, 4, 5, 10
    ds, obs, model, init = _experiment(N, k, n_cycles, S)
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, model_obj=model, ensemble_dim=k,
                           multiplicative_inflation=1.1)
    p = dc._plan(init, 0.0, obs, n_cycles, analysis_window=0.1)
    cfg = p['cfg']
    assert (cfg.system_dim, cfg.ensemble_dim, cfg.n_steps, cfg.n_cycles) == (N, k, S, n_cycles)
    assert cfg.rho == 1.1 and cfg.obs_error_sd == 0.5 and cfg.model_dt == 0.01 and cfg.stationary == 1
    assert p['vals'].shape == (n_cycles + 1, 3) and p['sysidx'].dtype == np.int64
    assert np.array_equal(p['padded'], np.arange(1, n_cycles + 1)[:, None])
    assert dc.steps_per_window == S + 1


def test_inline_model_wrapper_of_the_reference_tests_is_recognised():
    """The reference's cycler tests define their forecast plugin inline
    (tests/dacycler_etkf_test.py:39-46, tests/dacycler_var3d_test.py:36-45): a Model subclass around a
    Lorenz96 generator.  It is accepted as it is -- the plan takes delta_t and the forcing from the
    generator -- while a Model around anything else is refused."""
    N, k, n_cycles, S = 8, 4, 5, 10
    ds, obs, _, init = _experiment(N, k, n_cycles, S)

    class L96Model(dab.model.Model):
        def forecast(self, state_vec, n_steps):
            new_vec = self.model_obj.generate(x0=state_vec['x'].data, n_steps=n_steps)
            return new_vec.isel(time=-1), new_vec

    inline = L96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.02, forcing_term=7.5))
    p = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, model_obj=inline, ensemble_dim=k)._plan(
        init, 0.0, obs, n_cycles, analysis_window=0.1)
    assert p['cfg'].model_dt == 0.02 and p['cfg'].forcing == 7.5 and p['cfg'].n_steps == S

    class Other(dab.model.Model):
        def forecast(self, state_vec, n_steps):
            return state_vec, state_vec

    with pytest.raises(NotImplementedError):
        dab.dacycler.ETKF(system_dim=N, delta_t=0.01, model_obj=Other(model_obj=object()), ensemble_dim=k)._plan(
            init, 0.0, obs, n_cycles, analysis_window=0.1)
    with pytest.raises(NotImplementedError):
        dab.dacycler.ETKF(system_dim=N, delta_t=0.01, model_obj=dab.data.Lorenz96(system_dim=N), ensemble_dim=k)._plan(
            init, 0.0, obs, n_cycles, analysis_window=0.1)       # a bare generator is not a Model


def test_cycle_rejects_what_it_does_not_accelerate():
    N, k = 8, 4
    ds, obs, model, init = _experiment(N, k)
    with pytest.raises(AssertionError):                   # dabench/dacycler/_etkf.py:196-199
        dab.dacycler.ETKF(N, 0.01, model, ensemble_dim=k + 1)._plan(init, 0.0, obs, 5, analysis_window=0.1)
    with pytest.raises(NotImplementedError):
        dab.dacycler.ETKF(N, 0.01, model, ensemble_dim=k, H=np.eye(N))._plan(init, 0.0, obs, 5, analysis_window=0.1)
    with pytest.raises(ValueError):                       # dabench/dacycler/_dacycler.py:103-104
        dab.dacycler.ETKF(N, 0.01, model, ensemble_dim=k, h=lambda x: x)._plan(init, 0.0, obs, 5, analysis_window=0.1)

    class Other(dab.model.Model):
        def forecast(self, s, n_steps):
            return s, s
    with pytest.raises(NotImplementedError):
        dab.dacycler.ETKF(N, 0.01, Other(), ensemble_dim=k)._plan(init, 0.0, obs, 5, analysis_window=0.1)


def test_mini_dataset_surface():
    ds = _xr.MiniDataset({'x': (('time', 'index'), np.arange(12.0).reshape(3, 4))},
                         coords={'time': np.array([0.0, 0.1, 0.2]), 'index': np.arange(4)}, attrs={'a': 1})
    assert ds.sizes == {'time': 3, 'index': 4} and ds.a == 1
    last = ds.isel(time=-1)
    assert last['x'].dims == ('index',) and np.array_equal(last['x'].values, [8, 9, 10, 11])
    assert float(last['time'].data) == 0.2
    e = ds.assign(x=(('ensemble', 'index'), np.ones((2, 4))))
    assert e['x'].shape == (2, 4) and ds['x'].shape == (3, 4)
    assert np.array_equal(ds.mean(dim='time')['x'].values, [4, 5, 6, 7])
    s = _xr.MiniDataset({'x': (('cycle', 'ensemble', 'cycle_timestep', 'index'), np.zeros((2, 3, 5, 4)))})
    st = s.stack(time=['cycle', 'cycle_timestep']).transpose('time', ...)
    assert st['x'].shape == (10, 3, 4)


def test_cycle_batched_marshalling(monkeypatch):
    """cycle_batched hands the C ABI one observation set for a sweep (flag bits 0 and 1 of
    shared_index) and per-experiment arrays otherwise; no kernel runs here (fake handle)."""
    from dataassimbench_b200 import _runtime
    from dataassimbench_b200.dacycler import cycle_batched
    N, k, n_cycles, S = 8, 4, 5, 10
    ds, obs, model, init = _experiment(N, k, n_cycles, S)
    calls = []

    class Fake:
        def etkf_cycle_batched(self, cfg, batch, rho, x0, vals, idx, shared, padded, out, mean, fin):
            calls.append(dict(batch=batch, rho=np.array(rho), x0=x0.shape, vals=vals.shape, idx=idx.shape,
                              shared=shared, out=None if out is None else out.shape,
                              mean=None if mean is None else mean.shape))
            for a in (out, mean):
                if a is not None:
                    a[...] = 0.0
    monkeypatch.setattr(_runtime, 'handle', lambda: Fake())
    rhos = [1.0, 1.1, 1.3]
    cyclers = [dab.dacycler.ETKF(N, 0.01, model, ensemble_dim=k, multiplicative_inflation=r) for r in rhos]
    res = cycle_batched(cyclers, init, 0.0, obs, n_cycles, analysis_window=0.1)            # one obs set
    assert res['x'].shape == (3, n_cycles, k, N)
    c = calls[-1]
    assert c['shared'] == 3 and c['vals'] == (n_cycles + 1, 3) and c['idx'] == (n_cycles + 1, 3)
    assert c['batch'] == 3 and np.allclose(c['rho'], rhos) and c['x0'] == (3, k, N)
    res = cycle_batched(cyclers, [init] * 3, 0.0, [obs] * 3, n_cycles, analysis_window=0.1,
                        return_mean_only=True)                                              # a list of obs sets
    assert res['x_mean'].shape == (3, n_cycles, N)
    c = calls[-1]
    assert c['shared'] == 0 and c['vals'] == (3, n_cycles + 1, 3) and c['idx'] == (3, n_cycles + 1, 3)
    with pytest.raises(ValueError):                       # experiments of a batch must share their shape
        other = dab.dacycler.ETKF(N, 0.02, dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.02)),
                                  ensemble_dim=k)
        cycle_batched([cyclers[0], other], init, 0.0, obs, n_cycles, analysis_window=0.1)


def _var3d_reference_recipe():
    """The set-up of /root/reference/tests/dacycler_var3d_test.py:11-66 with this package's classes (the
    nature run comes from the oracle here: `generate()` needs the GPU)."""
    nature_x = o_l96.generate_dopri5(o_l96.default_x0(6, 0.05), 50, 0.05)
    times = o_l96.time_grid(50, 0.05)
    nature = _xr.new_dataset({'x': (('time', 'index'), nature_x)}, coords={'time': times, 'index': np.arange(6)})
    obs = dab.observer.Observer(nature, random_time_density=0.7, random_location_count=3, error_bias=0.0,
                                error_sd=0.7, random_seed=94, stationary_observers=True).observe()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=6, store_as_jax=True))
    dc = dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model)
    init = nature.isel(time=0)
    init = init.assign(x=(['index'], init['x'].values + threefry.normal(42, (6,))))
    return nature, obs, dc, init, times


class _FakeVar3DHandle:
    """The two device calls of Var3D.cycle answered by the oracle on CPU tensors."""

    def var3d_analysis(self, xb, xa, N, loc, val, n_rows, sd, stream=None):
        l = loc.numpy() if loc is not None else np.zeros(0, dtype=np.int64)
        v = val.numpy() if val is not None else np.zeros(0)
        assert len(l) == n_rows and xb.numel() == N
        xa.numpy()[:] = o_var3d.analysis_diagonal(xb.numpy(), v, l, np.full(len(l), 1.0 / sd ** 2))

    def l96_forecast(self, x_in, x_out, traj, N, k, n_steps, dt, F, stream=None):
        last, tr = o_l96.rk4_forecast(x_in.numpy(), n_steps, dt, F, return_traj=True)
        x_out.numpy()[:] = last
        if traj is not None:
            traj.numpy()[:n_steps] = tr[1:]


def test_var3d_cycle_host_logic(monkeypatch):
    """Var3D.cycle's bookkeeping (windows, valid rows, output assembly) against the oracle cycle, with the
    two device calls replaced by their oracle equivalents; the first analysis is the reference's presaved
    vector (tests/dacycler_var3d_test.py:70-77)."""
    pytest.importorskip("torch")
    nature, obs, dc, init, times = _var3d_reference_recipe()
    monkeypatch.setattr(dab.dacycler.Var3D, '_device', 'cpu')
    monkeypatch.setattr(dab.dacycler.Var3D, '_handle', lambda self: _FakeVar3DHandle())
    out = dc.cycle(input_state=init, start_time=times[0], obs_vector=obs, n_cycles=10, analysis_window=0.25,
                   return_forecast=False)
    assert out['x'].shape == (10, 6) and _xr.var_dims(out, 'x') == ('cycle', 'index')
    assert np.allclose(out['x'].values[0], [-0.90632236, -1.20861455, 1.64865068, 5.11034063, 4.399881, -3.75779771])
    fc = lambda x, n: o_l96.generate_rk4(x, n, 0.05)
    ref = o_var3d.cycle(init['x'].values, times[0], obs['x'].values, obs['time'].values,
                        obs['system_index'].values[0], 10, 0.7, fc, 0.05, analysis_window=0.25,
                        return_forecast=True, literal=True)
    assert np.max(np.abs(out['x'].values - ref[:, 0])) < 1e-9
    full = dc.cycle(input_state=init, start_time=times[0], obs_vector=obs, n_cycles=10, analysis_window=0.25,
                    return_forecast=True)
    assert full['x'].shape == (10, 5, 6) and _xr.var_dims(full, 'x') == ('cycle', 'cycle_timestep', 'index')
    assert np.max(np.abs(full['x'].values - ref)) < 1e-9
    assert np.array_equal(full['x'].values[:, 0], out['x'].values)


def test_var3d_rows_of_cycle_masks():
    """Two observation times in one window, moving observers with NaN padding, an empty window."""
    p = dict(padded=np.array([[1, 2], [3, 0], [0, 0]], dtype=np.int64), stationary=False,
             vals=np.array([[1.0, np.nan], [2.0, 3.0], [np.nan, np.nan]]),
             sysidx=np.array([[4, 0], [1, 4], [0, 0]], dtype=np.int64))
    loc, val = dab.dacycler.Var3D._rows_of_cycle(p, 0)
    assert loc.tolist() == [4, 1, 4] and val.tolist() == [1.0, 2.0, 3.0]
    loc, val = dab.dacycler.Var3D._rows_of_cycle(p, 1)
    assert len(loc) == 0 and len(val) == 0                    # its only slot is all padding
    loc, val = dab.dacycler.Var3D._rows_of_cycle(p, 2)
    assert len(loc) == 0
    p['stationary'] = True
    loc, val = dab.dacycler.Var3D._rows_of_cycle(p, 0)
    assert loc.tolist() == [4, 0, 1, 4] and np.isnan(val[1])  # stationary observers: all-true mask, as the reference


def test_var3d_rejects_what_it_does_not_accelerate():
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=6))
    nature, obs, dc, init, times = _var3d_reference_recipe()
    with pytest.raises(NotImplementedError):
        dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model, R=np.eye(3)).cycle(init, 0.0, obs, 2)
    with pytest.raises(ValueError, match='B must be'):          # a user B is accelerated (CG), a misshapen one is not
        dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model, B=np.eye(5))._plan(init, 0.0, obs, 2)
    assert dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model, B=2 * np.eye(6))._plan(init, 0.0, obs, 2)['B'][0, 0] == 2.0
    with pytest.raises(ValueError):
        dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model, h=lambda x: x).cycle(init, 0.0, obs, 2)
    class HostModel(dab.model.Model):                         # a user plugin without a device forecast: no fallback
        def forecast(self, state_vec, n_steps):
            return state_vec, state_vec
    with pytest.raises(NotImplementedError):
        dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=HostModel(system_dim=6, delta_t=0.05)).cycle(init, 0.0, obs, 2)
    with pytest.raises(NotImplementedError):
        dab.dacycler.ETKF(system_dim=6, delta_t=0.05, model_obj=HostModel(system_dim=6, delta_t=0.05), ensemble_dim=4)._plan(
            init.assign(x=(['ensemble', 'index'], np.zeros((4, 6)))), 0.0, obs, 2)


def test_compact_default_operators_match_the_reference_dense_ones():
    """DACycler._calc_default_H / _calc_default_R (dabench/dacycler/_dacycler.py:58-72) in their compact
    forms (index list, scaled identity): `dense()` is exactly the matrix the reference builds, masking is
    zeroing rows (dabench/dacycler/_etkf.py:202-203), and a dense one-hot matrix is recognised back."""
    from dataassimbench_b200.dacycler._dacycler import IndexOperator, ScaledIdentity, DACycler
    from oracle import etkf as o_etkf
    dc = DACycler(system_dim=9, delta_t=0.01, model_obj=None)
    idx = np.array([4, 0, 7, 7])
    H = dc._calc_default_H(np.zeros(4), idx)
    assert isinstance(H, IndexOperator) and H.shape == (4, 9)
    assert np.array_equal(H.dense(), o_etkf.calc_default_H(4, 9, idx))
    m = np.array([True, False, True, True])
    Hm = H.masked(m)
    assert np.array_equal(Hm.dense(), np.where(m, H.dense().T, 0).T)
    back = IndexOperator.from_dense(Hm.dense())
    assert np.array_equal(back.dense(), Hm.dense())
    with pytest.raises(NotImplementedError):
        IndexOperator.from_dense(np.full((2, 9), 0.5))
    R = dc._calc_default_R(np.zeros((1, 2, 2)), 1.5)
    assert isinstance(R, ScaledIdentity) and len(R) == 4
    assert np.array_equal(R.dense(), o_etkf.calc_default_R(4, 1.5))
    assert ScaledIdentity.sd_of(R) == 1.5 and ScaledIdentity.sd_of(R.dense()) == 1.5
    assert ScaledIdentity.sd_of(np.zeros((3, 3))) == 0.0
    # a diagonal R with unequal entries: one sigma per observation; off-diagonal entries are not accelerated
    assert np.array_equal(ScaledIdentity.sd_of(np.diag([1.0, 4.0])), [1.0, 2.0])
    assert np.array_equal(ScaledIdentity.sd_of(np.array([9.0, 4.0])), [3.0, 2.0])
    with pytest.raises(NotImplementedError):
        ScaledIdentity.sd_of(np.array([[1.0, 0.1], [0.1, 1.0]]))
    with pytest.raises(NotImplementedError):
        ScaledIdentity.sd_of(np.diag([1.0, -1.0]))
    assert len(dc._calc_default_R(np.zeros(0), 1.0)) == 0          # the empty-R branch tests len(R)


def test_etkf_grid_sharding_switch_without_a_process_group():
    """`ETKF(distributed=...)`: no torch.distributed group -> single GPU unless sharding was demanded."""
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=4096, delta_t=0.01))
    assert dab.dacycler.ETKF(system_dim=4096, delta_t=0.01, model_obj=model)._shard_group(4096, 20) is None
    assert dab.dacycler.ETKF(system_dim=4096, delta_t=0.01, model_obj=model, distributed=False)._shard_group(4096, 20) is None
    with pytest.raises(ValueError, match='distributed=True needs'):
        dab.dacycler.ETKF(system_dim=4096, delta_t=0.01, model_obj=model, distributed=True)._shard_group(4096, 20)
