"""GPU tests through the reference-facing Python API (`ETKF.cycle()` etc.) and the golden fixtures."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dataassimbench_b200 as dab
from dataassimbench_b200 import _xr, _native
from oracle import l96 as o_l96, etkf as o_etkf, var3d as o_var3d, threefry, metrics as o_metrics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_reference_etkf_test_recipe_through_the_api():
    """The reference's own test (tests/dacycler_etkf_test.py:13-102), written against this package.
    The forecast is fixed-step RK4 at the model's dt = 0.05 instead of adaptive Dormand-Prince, so
    only the first-cycle golden (:89-92, a pure analysis) can be asserted literally; the rest of the
    run is compared pointwise with the oracle cycling with the SAME integrator."""
    l96 = dab.data.Lorenz96(system_dim=5, store_as_jax=True, delta_t=0.01)
    nature = l96.generate(n_steps=25)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, 25, 5)],
                                random_location_count=3, error_bias=0.1, error_sd=1.0, random_seed=91,
                                stationary_observers=True, store_as_jax=True).observe()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=5, store_as_jax=True, delta_t=0.05))
    dc = dab.dacycler.ETKF(system_dim=5, delta_t=0.01, ensemble_dim=8, model_obj=model)
    init = nature.isel(time=10)
    noise = threefry.normal(42, (8, 5))
    init = init.assign(x=(['ensemble', 'index'], init['x'].values + noise))
    start_time = init['time'].data
    out_sv = dc.cycle(input_state=init, start_time=start_time, obs_vector=obs, obs_error_sd=1.5,
                      analysis_window=0.1, n_cycles=10, return_forecast=True)
    assert out_sv['x'].shape == (10, 8, 10, 5)
    flat = out_sv.stack(time=['cycle', 'cycle_timestep']).transpose('time', ...)
    assert flat['x'].shape == (100, 8, 5)
    assert flat.mean(dim='ensemble')['x'].shape == (100, 5)
    v = flat['x'].values
    assert not np.allclose(v[-1, 1, :], v[-1, 0, :])
    assert np.allclose(v[0, 0, :], [-0.85402591, 1.03480315, 0.51005132, 6.61546551, 8.1166806])
    # same experiment in the oracle with RK4 at dt = 0.05
    fc = lambda x, n: o_l96.generate_rk4(x, n, 0.05)
    ref = o_etkf.cycle(init['x'].values, float(start_time), obs['x'].values, obs['time'].values,
                       obs['system_index'].values[0], 10, 1.5, fc, 0.01, analysis_window=0.1,
                       return_forecast=True, literal=True)
    assert rel_err(out_sv['x'].values, ref) < 1e-9
    # analyses only
    out_a = dc.cycle(input_state=init, start_time=start_time, obs_vector=obs, obs_error_sd=1.5,
                     analysis_window=0.1, n_cycles=10)
    assert out_a['x'].shape == (10, 8, 5)
    assert np.array_equal(out_a['x'].values, out_sv['x'].values[:, :, 0, :])
    # trajectories streamed to the host in blocks of cycles (3, 3, 3, 1 here) are the same array
    dc.traj_block_bytes = 3 * (10 * 8 * 5 * 8) + 17
    out_blk = dc.cycle(input_state=init, start_time=start_time, obs_vector=obs, obs_error_sd=1.5,
                       analysis_window=0.1, n_cycles=10, return_forecast=True)
    assert np.array_equal(out_blk['x'].values, out_sv['x'].values)


def test_generate_and_model_forecast_contract():
    """Model.forecast -> (last_state, trajectory with leading 'time' of length n_steps),
    SURVEY.md 3.2 / tests/model_base_test.py."""
    gen = dab.data.Lorenz96(system_dim=36, delta_t=0.01)
    traj = gen.generate(n_steps=21)
    assert traj['x'].shape == (21, 36) and traj.attrs['delta_t'] == 0.01
    assert np.allclose(traj['time'].values, np.arange(21) * 0.01)
    assert rel_err(traj['x'].values, o_l96.generate_rk4(gen.x0, 21, 0.01)) < 1e-12
    t2 = gen.generate(t_final=0.1)
    assert t2['x'].shape == (10, 36)
    model = dab.model.Lorenz96Model(model_obj=gen)
    last, tr = model.forecast(traj.isel(time=3), n_steps=5)
    assert tr['x'].shape == (5, 36) and last['x'].shape == (36,)
    assert np.array_equal(last['x'].values, tr['x'].values[-1])
    assert np.array_equal(tr['x'].values[0], traj['x'].values[3])


def test_lorenz63_model_forecast_reference_known_answer():
    """/root/reference/tests/model_base_test.py:18-30 and tests/data_lorenz63_test.py:22-75 through
    the API classes (RK4 at delta_t instead of the adaptive integrator: within the tests' allclose)."""
    m = dab.model.Lorenz63Model(model_obj=dab.data.Lorenz63())
    sv = _xr.new_dataset({'x': (('index',), m.model_obj.x0)})
    new, traj = m.forecast(sv)
    assert np.allclose(sv['x'].values, np.array([-10., -15., 21.3]))
    assert not np.allclose(new['x'].values, sv['x'].values)
    assert np.allclose(new['x'].values, np.array([-10.499835, -15.48437, 22.282299]))
    assert traj['x'].shape == (2, 3)
    g = dab.data.Lorenz63(system_dimension=3)
    t1 = g.generate(t_final=1)
    assert t1['x'].shape == (100, 3) and t1.attrs['system_dim'] == 3
    assert np.array_equal(t1['x'].values, dab.data.Lorenz63().generate(t_final=1)['x'].values)
    t2 = g.generate(t_final=1, x0=np.array([-2.2, -2.2, 19.1]))
    assert not np.allclose(t1['x'].values, t2['x'].values, rtol=1e-5, atol=0)
    assert not np.allclose(t2['x'].values[-1], np.array([-2.2, -2.2, 19.1]))
    from oracle import l63 as o_l63
    assert rel_err(t1['x'].values, o_l63.generate_rk4(o_l63.X0_DEFAULT, 100, 0.01)) < 1e-11
    # against the reference's own integrator the fixed-step RK4 differs by its truncation error only
    assert rel_err(t1['x'].values[:30], o_l63.generate_dopri5(o_l63.X0_DEFAULT, 30, 0.01)) < 1e-5


def test_obsop_gather_is_index_selection():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 50))
    idx = rng.choice(50, 17, replace=False)
    assert np.array_equal(dab.obsop.ObsOp.gather(X, idx), X[:, idx])
    op = dab.obsop.ObsOp(location_indices=idx)
    assert np.array_equal(op.h(dab.vector.StateVector(values=X[0])), X[0, idx])
    with pytest.raises(IndexError):
        dab.obsop.ObsOp.gather(X, [50])


@pytest.mark.parametrize("stationary", [True, False])
@pytest.mark.parametrize("N,n_times", [(40, 25), (5000, 12)])
def test_observer_samples_a_device_trajectory_bit_for_bit(stationary, N, n_times):
    """Observer.observe with the trajectory in HBM (dab_observe_f64) = the same call on the host copy:
    same draws, same values to the bit, NaN in the ragged tails (observer/_observer.py:290-316, 346-363)."""
    rng = np.random.default_rng(N + n_times)
    traj = rng.standard_normal((n_times, N))
    coords = {'time': np.arange(n_times) * 0.05, 'index': np.arange(N)}
    kw = dict(random_time_density=0.5, random_location_density=0.3, error_sd=0.7, error_bias=0.05,
              stationary_observers=stationary, random_seed=5)
    host = dab.observer.Observer(_xr.new_dataset({'x': (('time', 'index'), traj)}, coords=coords), **kw).observe()
    dev = dab.observer.Observer(_xr.new_dataset({'x': (('time', 'index'), torch.from_numpy(traj).cuda())},
                                                coords=coords), **kw).observe()
    for name in ('x', 'system_index', 'errors'):
        a, b = np.asarray(host[name].values), np.asarray(dev[name].values)
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), name
    assert np.array_equal(host['time'].values, dev['time'].values)
    if not stationary:
        assert np.isnan(np.asarray(dev['x'].values)).any()          # ragged rows were padded
    # the reference's relation between the three fields
    v, e, si = (np.asarray(dev[n].values) for n in ('x', 'errors', 'system_index'))
    pos = np.searchsorted(coords['time'], dev['time'].values)
    ok = ~np.isnan(v)
    assert np.array_equal(v[ok], (traj[pos][np.arange(len(pos))[:, None], si[0]] + e[0])[ok])


@pytest.mark.parametrize("stationary", [True, False])
def test_observer_device_route_matches_the_oracle_observer(stationary):
    """The device route of Observer.observe (draws on the host, sampling by dab_observe_f64) directly against
    oracle/observer.py (the restatement pinned by the reference's own known answers,
    tests/observer_base_test.py:133-180): indices, times, errors and values to the bit."""
    from oracle import observer as o_obs
    rng = np.random.default_rng(3)
    N, n_times = 3000, 20
    traj = rng.standard_normal((n_times, N))
    times = np.arange(n_times) * 0.05
    kw = dict(random_time_density=0.6, random_location_density=0.25, error_sd=0.9, error_bias=-0.1,
              stationary_observers=stationary, random_seed=17)
    ref = o_obs.observe(traj, times, **kw)
    ds = dab.observer.Observer(_xr.new_dataset({'x': (('time', 'index'), torch.from_numpy(traj).cuda())},
                                               coords={'time': times, 'index': np.arange(N)}), **kw).observe()
    assert np.array_equal(np.asarray(ds['time'].values), ref['time'])
    assert np.array_equal(np.asarray(ds['system_index'].values)[0], ref['system_index'])
    assert np.array_equal(np.asarray(ds['errors'].values)[0], ref['errors'], equal_nan=True)
    assert np.array_equal(np.asarray(ds['x'].values), ref['values'], equal_nan=True)


def test_observe_entry_point_rejects_bad_indices():
    from dataassimbench_b200._runtime import handle
    H = handle()
    state = torch.arange(12, dtype=torch.float64, device='cuda').reshape(3, 4)
    out = torch.empty((2, 3), dtype=torch.float64, device='cuda')
    pos = torch.tensor([0, 2], dtype=torch.int64, device='cuda')
    loc = torch.tensor([3, 0, 1], dtype=torch.int64, device='cuda')
    H.observe(state, 4, 3, pos, 2, loc, True, None, 3, None, out)
    assert np.array_equal(out.cpu().numpy(), np.array([[3., 0., 1.], [11., 8., 9.]]))
    cnt = torch.tensor([1, 3], dtype=torch.int64, device='cuda')
    loc2 = torch.tensor([[2, 99, 99], [1, 1, 0]], dtype=torch.int64, device='cuda')      # 99 sits in a padded slot
    H.observe(state, 4, 3, pos, 2, loc2, False, cnt, 3, None, out)
    assert np.array_equal(out.cpu().numpy(), np.array([[2., np.nan, np.nan], [9., 9., 8.]]), equal_nan=True)
    with pytest.raises(RuntimeError, match="out of range"):
        H.observe(state, 4, 3, pos, 2, torch.tensor([0, 4, 1], dtype=torch.int64, device='cuda'), True, None, 3,
                  None, out)
    with pytest.raises(RuntimeError, match="out of range"):
        H.observe(state, 4, 3, torch.tensor([0, 3], dtype=torch.int64, device='cuda'), 2, loc, True, None, 3,
                  None, out)


def test_var3d_analysis_entry_point():
    """dab_var3d_analysis_f64 = the closed form of the reference's CG system for default H / R / B
    (oracle/var3d.py::analysis_diagonal, itself checked against the literal CG solve on CPU)."""
    from dataassimbench_b200._runtime import handle
    H = handle()
    rng = np.random.default_rng(8)
    N, sd = 5000, 0.7
    xb = rng.standard_normal(N)
    d_xb = torch.from_numpy(xb).cuda()
    d_xa = torch.empty_like(d_xb)
    # one observation per point: bit-exact
    loc = rng.choice(N, 1500, replace=False, shuffle=False).astype(np.int64)
    val = rng.standard_normal(1500)
    H.var3d_analysis(d_xb, d_xa, N, torch.from_numpy(loc).cuda(), torch.from_numpy(val).cuda(), 1500, sd)
    ref = o_var3d.analysis_diagonal(xb, val, loc, np.full(1500, 1.0 / sd ** 2))
    assert np.array_equal(d_xa.cpu().numpy(), ref)
    assert not np.array_equal(ref, xb) and np.array_equal(ref[np.setdiff1d(np.arange(N), loc)], xb[np.setdiff1d(np.arange(N), loc)])
    # two observation times in the window hitting the same points, plus a triple: sums in any order
    loc2 = np.concatenate([loc, loc[:700], loc[:5]])
    val2 = np.concatenate([val, rng.standard_normal(705)])
    H.var3d_analysis(d_xb, d_xa, N, torch.from_numpy(loc2).cuda(), torch.from_numpy(val2).cuda(), len(loc2), sd)
    ref2 = o_var3d.analysis_diagonal(xb, val2, loc2, np.full(len(loc2), 1.0 / sd ** 2))
    assert np.max(np.abs(d_xa.cpu().numpy() - ref2)) < 1e-14
    # no observation: the analysis is the background
    H.var3d_analysis(d_xb, d_xa, N, None, None, 0, sd)
    assert np.array_equal(d_xa.cpu().numpy(), xb)
    with pytest.raises(RuntimeError, match="out of range"):
        H.var3d_analysis(d_xb, d_xa, N, torch.tensor([0, N], dtype=torch.int64, device='cuda'),
                         torch.zeros(2, dtype=torch.float64, device='cuda'), 2, sd)


def test_var3d_reference_recipe_through_the_api():
    """The reference's Var3D test (tests/dacycler_var3d_test.py:11-84) written against this package: the first
    analysis is its presaved vector; the run equals the oracle's literal cycle (dense H/R/B + conjugate
    gradients) with the same fixed-step RK4 forecast."""
    nature = dab.data.Lorenz96(system_dim=6, store_as_jax=True).generate(n_steps=50)
    obs = dab.observer.Observer(nature, random_time_density=0.7, random_location_count=3, error_bias=0.0,
                                error_sd=0.7, random_seed=94, stationary_observers=True).observe()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=6, store_as_jax=True))
    dc = dab.dacycler.Var3D(system_dim=6, delta_t=0.05, model_obj=model)
    init = nature.isel(time=0)
    x_init = init['x'].values + threefry.normal(42, (6,))
    init = init.assign(x=(['index'], x_init.copy()))
    t0 = nature['time'].values[0]
    out = dc.cycle(input_state=init, start_time=t0, obs_vector=obs, n_cycles=10, analysis_window=0.25,
                   return_forecast=False)
    assert out['x'].shape == (10, 6)
    assert np.allclose(out['x'].values[0], [-0.90632236, -1.20861455, 1.64865068, 5.11034063, 4.399881, -3.75779771])
    assert np.array_equal(init['x'].values, x_init)                       # the caller's state is untouched
    fc = lambda x, n: o_l96.generate_rk4(x, n, 0.05)
    ref = o_var3d.cycle(x_init, t0, obs['x'].values, obs['time'].values, obs['system_index'].values[0], 10, 0.7,
                        fc, 0.05, analysis_window=0.25, return_forecast=True, literal=True)
    assert rel_err(out['x'].values, ref[:, 0]) < 1e-9
    full = dc.cycle(input_state=init, start_time=t0, obs_vector=obs, n_cycles=10, analysis_window=0.25,
                    return_forecast=True)
    assert full['x'].shape == (10, 5, 6)
    assert rel_err(full['x'].values, ref) < 1e-9
    assert np.array_equal(full['x'].values[:, 0], out['x'].values)


@pytest.mark.parametrize("tag", ["c1", "c2", "c3s", "c5s"])
def test_analysis_golden_fixture(tag):
    """Committed literal-formula vectors (tests/golden/make_golden.py): <= 1e-10."""
    z = np.load(os.path.join(GOLD, "analysis_cases.npz"))
    X, idx, y, mask, Xa = (z["%s_%s" % (tag, n)] for n in ("X", "idx", "y", "mask", "Xa"))
    rho, sd = (float(v) for v in z[tag + "_par"])
    k, N = X.shape
    H = _native.Handle(0)
    out = torch.empty((k, N), dtype=torch.float64, device='cuda')
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    H.etkf_analysis(c(X), out, c(y), c(idx), c(mask), sd, rho, N, k, len(idx),
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), Xa) < 1e-10
    H.close()


def test_multicycle_rmse_within_one_percent_of_reference_style_run():
    """P4 (north star: multi-cycle runs agree within 1 % on time-mean analysis RMSE).  Converged
    configuration: N=40, k=20, window 0.05, 40 obs, rho=1.05, 200 cycles.  The
    fixture holds the oracle run with the LITERAL analysis and the reference's adaptive
    Dormand-Prince forecast; the GPU run uses fixed-step RK4."""
    z = np.load(os.path.join(GOLD, "etkf_converged_dp5.npz"))
    N, k, S, dt, n_cycles, sd, rho = z["params"]
    N, k, S, n_cycles = int(N), int(k), int(S), int(n_cycles)
    padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=rho,
                              obs_error_sd=sd, n_obs_times=z["vals"].shape[0], n_obs=z["vals"].shape[1],
                              stationary=1, n_cycles=n_cycles, t_pad=1, rank=0, world=1)
    H = _native.Handle(0)
    out = np.empty((n_cycles, k, N))
    H.etkf_cycle_host(cfg, np.ascontiguousarray(z["X0"]), np.ascontiguousarray(z["vals"]),
                      np.ascontiguousarray(z["sysidx"]), padded, out)
    H.close()
    rmse = np.array([dab.metrics.rmse(out[c].mean(axis=0), z["truth"][c]) for c in range(n_cycles)])
    skip = n_cycles // 4
    gpu, ref = rmse[skip:].mean(), z["rmse_dp5"][skip:].mean()
    assert abs(gpu - ref) / ref < 0.01, (gpu, ref)
    assert dab.metrics.rmse(out[-1].mean(axis=0), z["truth"][-1]) == pytest.approx(
        o_metrics.rmse(out[-1].mean(axis=0), z["truth"][-1]))


def test_cycle_batched_inflation_sweep_matches_single_cycles():
    """dacycler.cycle_batched: same objects as cycle(), one launch for the whole sweep."""
    N, k, n_cycles, S = 40, 20, 10, 20
    gen = dab.data.Lorenz96(system_dim=N, delta_t=0.01)
    nature = gen.generate(n_steps=n_cycles * S + 1)
    obs = dab.observer.Observer(nature, times=nature['time'].values[::S], random_location_count=20,
                                error_sd=1.0, random_seed=5).observe()
    model = dab.model.Lorenz96Model(model_obj=gen)
    X0 = nature['x'].values[0] + np.random.default_rng(1).standard_normal((k, N))
    init = _xr.new_dataset({'x': (('ensemble', 'index'), X0)}, coords={'index': np.arange(N)})
    rhos = [1.0, 1.1, 1.25, 1.4]
    cyclers = [dab.dacycler.ETKF(system_dim=N, delta_t=0.01, model_obj=model, ensemble_dim=k,
                                 multiplicative_inflation=r) for r in rhos]
    out = dab.dacycler.cycle_batched(cyclers, init, 0.0, obs, n_cycles, analysis_window=0.2)
    assert out['x'].shape == (4, n_cycles, k, N)
    for b, dc in enumerate(cyclers):
        single = dc.cycle(init, 0.0, obs, n_cycles, analysis_window=0.2)
        assert rel_err(out['x'].values[b], single['x'].values) < 1e-10
    m = dab.dacycler.cycle_batched(cyclers, init, 0.0, obs, n_cycles, analysis_window=0.2, return_mean_only=True)
    assert rel_err(m['x_mean'].values, out['x'].values.mean(axis=2)) < 1e-12


# ------------------------------------------------------------------ the reference's overridable hooks
def _hook_case(N=120, k=12, n_obs=40, seed=3):
    rng = np.random.default_rng(seed)
    Xb = 8.0 + rng.standard_normal((N, k))                      # reference orientation: (system_dim, ensemble_dim)
    idx = rng.choice(N, n_obs, replace=False, shuffle=False)
    y = Xb.mean(axis=1)[idx] + 0.7 * rng.standard_normal(n_obs)
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01))
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model,
                           multiplicative_inflation=1.1)
    return dc, Xb, idx, y


def test_hook_apply_obsop_is_the_reference_H_at_Xb_bit_for_bit():
    """ETKF._apply_obsop (dabench/dacycler/_etkf.py:82-92) with the default H, against the dense
    one-hot product the reference forms (oracle: default_H @ Xb is a pure selection, so bit-exact)."""
    dc, Xb, idx, y = _hook_case()
    H = dc._calc_default_H(y, idx)
    Yb = dc._apply_obsop(Xb, H, None)
    assert Yb.shape == (len(idx), Xb.shape[1])
    assert np.array_equal(Yb, o_etkf.calc_default_H(len(idx), Xb.shape[0], idx) @ Xb)
    assert np.array_equal(Yb, Xb[idx])
    # the dense matrix the reference would pass works too (it is recognised as an index gather) ...
    assert np.array_equal(dc._apply_obsop(Xb, H.dense(), None), Yb)
    # ... masked rows are zero rows of H
    m = np.arange(len(idx)) % 3 != 0
    assert np.array_equal(dc._apply_obsop(Xb, H.masked(m), None), np.where(m[:, None], Xb[idx], 0.0))
    with pytest.raises(NotImplementedError):
        dc._apply_obsop(Xb, np.ones((3, Xb.shape[0])), None)
    with pytest.raises(ValueError):
        dc._apply_obsop(Xb, None, lambda x: x)


def test_hook_compute_analysis_matches_the_literal_formula():
    """ETKF._compute_analysis(Xb, Y, H, h, R, rho) (dabench/dacycler/_etkf.py:94-163), same signature and
    orientation, against the oracle's literal restatement (dense H, dense R, pinv, sqrtm): <= 1e-10."""
    dc, Xb, idx, y = _hook_case()
    N, k = Xb.shape
    H = dc._calc_default_H(y, idx)
    R = dc._calc_default_R(y, 0.7)
    Xa = dc._compute_analysis(Xb, y, H, None, R, rho=1.1)
    ref = o_etkf.analysis_literal(Xb, y, o_etkf.calc_default_H(len(idx), N, idx), o_etkf.calc_default_R(len(idx), 0.7), 1.1)
    assert Xa.shape == (N, k) and rel_err(Xa, ref) < 1e-10
    # dense operands as the reference builds them
    Xa2 = dc._compute_analysis(Xb, y.reshape(1, -1), H.dense(), None, R.dense(), rho=1.1)
    assert np.array_equal(Xa2, Xa)
    # empty R: the reference's zero branch (:149-152) collapses the ensemble onto its mean
    Xe = dc._compute_analysis(Xb, np.zeros(0), None, None, np.zeros((0, 0)), rho=1.1)
    ref_e = o_etkf.analysis_literal(Xb, np.zeros(0), np.zeros((0, N)), np.zeros((0, 0)), 1.1)
    assert rel_err(Xe, ref_e) < 1e-12
    # a diagonal R with unequal entries is accelerated; one with off-diagonal entries is refused, not approximated
    Rd = np.diag(np.linspace(1, 2, len(idx)))
    assert rel_err(dc._compute_analysis(Xb, y, H, None, Rd, 1.0), o_etkf.analysis_literal(Xb, y, H.dense(), Rd, 1.0)) < 1e-10
    with pytest.raises(NotImplementedError):
        dc._compute_analysis(Xb, y, H, None, Rd + 0.01, 1.0)


def test_hook_cycle_obsop_applies_masks_and_defaults():
    """ETKF._cycle_obsop (dabench/dacycler/_etkf.py:165-213): defaults resolved, time and location masks
    zero rows of H, Dataset in, Dataset `x[ensemble, i]` out."""
    dc, Xb, idx, y = _hook_case()
    N, k = Xb.shape
    dc.obs_error_sd = 0.7
    tmask = np.arange(len(idx)) % 4 != 1
    lmask = np.arange(len(idx)) % 5 != 2
    ds = _xr.new_dataset({'x': (('ensemble', 'index'), np.ascontiguousarray(Xb.T))})
    out = dc._cycle_obsop(ds, y.reshape(1, 1, -1), idx, tmask, lmask)
    assert tuple(out['x'].dims) == ('ensemble', 'i')
    Hd = o_etkf.calc_default_H(len(idx), N, idx)
    Hd = np.where(tmask, Hd.T, 0).T
    Hd = np.where(lmask, Hd.T, 0).T
    ref = o_etkf.analysis_literal(Xb, y, Hd, o_etkf.calc_default_R(len(idx), 0.7), 1.1)
    assert rel_err(out['x'].values.T, ref) < 1e-10
    bad = _xr.new_dataset({'x': (('ensemble', 'index'), np.ascontiguousarray(Xb.T[:-1]))})
    with pytest.raises(AssertionError):
        dc._cycle_obsop(bad, y, idx, tmask, lmask)


def test_hook_step_forecast_is_the_member_loop_of_the_reference():
    """ETKF._step_forecast(Xa, n_steps) (dabench/dacycler/_etkf.py:64-80): (last, trajectories) with the
    reference's dims; trajectory point 0 is Xa itself, `n_steps` points, against oracle RK4."""
    dc, Xb, idx, y = _hook_case()
    Xa = np.ascontiguousarray(Xb.T)
    ds = _xr.new_dataset({'x': (('ensemble', 'index'), Xa)})
    last, traj = dc._step_forecast(ds, n_steps=6)
    assert last['x'].shape == Xa.shape and traj['x'].shape == (Xa.shape[0], 6, Xa.shape[1])
    assert tuple(traj['x'].dims) == ('ensemble', 'time', 'index')
    assert np.array_equal(traj['x'].values[:, 0], Xa)
    assert np.array_equal(traj['x'].values[:, -1], last['x'].values)
    ref = np.stack([o_l96.generate_rk4(Xa[j], 6, 0.01) for j in range(Xa.shape[0])])
    assert rel_err(traj['x'].values, ref) < 1e-12
    # n_steps = 1: nothing to integrate
    last1, traj1 = dc._step_forecast(ds, n_steps=1)
    assert np.array_equal(last1['x'].values, Xa) and traj1['x'].shape == (Xa.shape[0], 1, Xa.shape[1])


def test_cycle_with_zero_error_sd_ignores_the_observations_like_pinv_of_zero_R():
    """Observer's default error_sd is 0.0 and cycle() falls back to it: R = 0, pinv(R) = 0
    (dabench/dacycler/_etkf.py:142), so the analysis is the inflated background -- finite, not NaN."""
    N, k, S = 64, 6, 4
    gen = dab.data.Lorenz96(system_dim=N, delta_t=0.01)
    nature = gen.generate(n_steps=40)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, 40, 4)], random_location_count=20,
                                random_seed=5, stationary_observers=True).observe()      # error_sd = 0.0
    assert obs.attrs['error_sd'] == 0.0
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k,
                           model_obj=dab.model.Lorenz96Model(model_obj=gen), multiplicative_inflation=1.21)
    rng = np.random.default_rng(2)
    X0 = nature['x'].values[0] + rng.standard_normal((k, N))
    init = _xr.new_dataset({'x': (('ensemble', 'index'), X0)})
    out = dc.cycle(init, 0.0, obs, n_cycles=3, analysis_window=S * 0.01)
    assert np.isfinite(out['x'].values).all()
    # literal oracle with R = 0
    Xb = X0.T
    ref = o_etkf.analysis_literal(Xb, obs['x'].values[0], o_etkf.calc_default_H(20, N, obs["system_index"].values[0, 0]),
                                  np.zeros((20, 20)), 1.21)
    assert rel_err(out['x'].values[0].T, ref) < 1e-10


def test_diagonal_R_with_unequal_entries_single_analysis_and_cycle():
    """A user R = diag(sd_r^2) (the reference takes pinv(R), dabench/dacycler/_etkf.py:142) through
    `_compute_analysis` against the LITERAL formula, and a vector obs_error_sd through `cycle()` against the
    oracle cycling with the same per-observation weights; sigma = 0 rows carry no weight (pinv)."""
    rng = np.random.default_rng(8)
    N, k, n_y = 60, 12, 25
    X = 2.0 + rng.standard_normal((N, k))
    idx = rng.choice(N, n_y, replace=False)
    y = rng.standard_normal(n_y) + 2.0
    sd = 0.5 + rng.random(n_y)
    sd[3] = 0.0
    R = np.diag(sd ** 2)
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01))
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model)
    Hm = o_etkf.calc_default_H(n_y, N, idx)
    ref = o_etkf.analysis_literal(X, y, Hm, R, 1.1)
    got = dc._compute_analysis(Xb=X, Y=y, H=Hm, h=None, R=R, rho=1.1)
    assert rel_err(got, ref) < 1e-10
    # through cycle(): 5 cycles, stationary network, one sigma per observation
    N, k, n_cycles, S = 256, 10, 5, 20
    gen = dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True)
    nature = gen.generate(n_steps=n_cycles * S + 1)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, n_cycles * S + 1, S)],
                                random_location_count=80, error_sd=1.0, random_seed=3, stationary_observers=True,
                                store_as_jax=True).observe()
    sdv = 0.5 + rng.random(80)
    sdv[7] = 0.0
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True))
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model)
    init = nature.isel(time=0)
    X0 = init['x'].values + rng.standard_normal((k, N))
    init = init.assign(x=(['ensemble', 'index'], X0))
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=n_cycles, obs_error_sd=sdv,
                   analysis_window=0.2, analysis_time_in_window=0.0)
    w = np.where(sdv == 0.0, 0.0, 1.0 / np.where(sdv == 0.0, 1.0, sdv) ** 2)
    Xc = X0.copy()
    loc = obs['system_index'].values[0][0]
    for c in range(n_cycles):
        Xa = o_etkf.analysis_sparse(Xc, obs['x'].values[c], loc, w, 1.0)
        assert rel_err(out['x'].values[c], Xa) < 1e-9, c
        Xc = o_l96.rk4_forecast(Xa, S, 0.01)
    # the same R handed over as a matrix on the cycler
    dc2 = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model, R=np.diag(sdv ** 2))
    out2 = dc2.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=n_cycles, obs_error_sd=1.0,
                     analysis_window=0.2, analysis_time_in_window=0.0)
    assert np.array_equal(out2['x'].values, out['x'].values)


def test_var3d_with_a_user_B_runs_the_reference_conjugate_gradients_on_the_device():
    """Var3D with a dense user B (dabench/dacycler/_var3d.py:84-106): `dab_var3d_analysis_b_f64` against the
    oracle's literal restatement (dense H, R, B; jax-style cg from x0 = xb, tol 1e-5) -- same iteration count, same
    iterate -- and a short cycle through the API."""
    rng = np.random.default_rng(21)
    N, n_y, sd = 96, 40, 0.8
    xb = 3.0 + rng.standard_normal(N)
    idx = rng.choice(N, n_y, replace=True).astype(np.int64)       # some points observed twice
    y = xb[idx] + rng.standard_normal(n_y)
    d = np.abs(np.arange(N)[:, None] - np.arange(N)[None, :])
    d = np.minimum(d, N - d)
    B = 0.6 * np.exp(-(d / 3.0) ** 2) + 0.4 * np.identity(N)       # SPD, periodic correlations
    Hm = o_etkf.calc_default_H(n_y, N, idx)
    A = np.identity(N) + (B @ Hm.T) @ np.linalg.inv(o_etkf.calc_default_R(n_y, sd)) @ Hm
    b1 = xb + (B @ Hm.T) @ np.linalg.inv(o_etkf.calc_default_R(n_y, sd)) @ y
    ref, ref_it = o_var3d.cg(A, b1, xb, tol=1e-5, maxiter=1000)
    assert np.allclose(o_var3d.analysis_literal(xb, y, Hm, o_etkf.calc_default_R(n_y, sd), B), ref, rtol=0, atol=1e-13)
    from dataassimbench_b200._runtime import handle
    xa = torch.empty(N, dtype=torch.float64, device='cuda')
    it = handle().var3d_analysis_b(torch.from_numpy(xb).cuda(), xa, N, torch.from_numpy(idx).cuda(),
                                   torch.from_numpy(y).cuda(), n_y, sd, torch.from_numpy(B).cuda(), 1e-5, 1000,
                                   torch.cuda.current_stream().cuda_stream)
    assert it == ref_it and it > 1
    assert rel_err(xa.cpu().numpy(), ref) < 1e-11
    # tighter tolerance converges to the direct solution
    it2 = handle().var3d_analysis_b(torch.from_numpy(xb).cuda(), xa, N, torch.from_numpy(idx).cuda(),
                                    torch.from_numpy(y).cuda(), n_y, sd, torch.from_numpy(B).cuda(), 1e-13, 1000,
                                    torch.cuda.current_stream().cuda_stream)
    assert it2 >= it and rel_err(xa.cpu().numpy(), np.linalg.solve(A, b1)) < 1e-10
    # through Var3D.cycle(): first analysis = the literal oracle's
    Nn = 40
    gen = dab.data.Lorenz96(system_dim=Nn, delta_t=0.01, store_as_jax=True)
    nature = gen.generate(n_steps=61)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, 61, 20)], random_location_count=20,
                                error_sd=0.5, random_seed=4, stationary_observers=True, store_as_jax=True).observe()
    dd = np.abs(np.arange(Nn)[:, None] - np.arange(Nn)[None, :])
    dd = np.minimum(dd, Nn - dd)
    Bn = 0.5 * np.exp(-(dd / 2.0) ** 2) + 0.5 * np.identity(Nn)
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=Nn, delta_t=0.01, store_as_jax=True))
    dc = dab.dacycler.Var3D(system_dim=Nn, delta_t=0.01, model_obj=model, B=Bn)
    init = nature.isel(time=0)
    x0 = init['x'].values + 0.5 * rng.standard_normal(Nn)
    init = init.assign(x=(['index'], x0))
    # (3 cycles x 0.2 trips the reference's float-arange quirk that _windows.py preserves: 2 cycles)
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=2, obs_error_sd=0.5,
                   analysis_window=0.2, analysis_time_in_window=0.0)
    loc = obs['system_index'].values[0][0]
    Hn = o_etkf.calc_default_H(20, Nn, loc)
    first = o_var3d.analysis_literal(x0, obs['x'].values[0], Hn, o_etkf.calc_default_R(20, 0.5), Bn)
    assert rel_err(out['x'].values[0], first) < 1e-11
    assert len(dc.cg_iterations) == 2 and all(i > 0 for i in dc.cg_iterations)
    assert np.all(np.isfinite(out['x'].values))


@pytest.mark.parametrize("stationary", [True, False])
def test_device_resident_observations_go_straight_into_the_cycler(stationary):
    """Observer(keep_on_device=True) on a trajectory in HBM leaves the observed values there and ETKF.cycle() packs
    them into its location-ordered tables with a kernel (dab_cycler_setup_dev): identical analyses to the same
    observations taken through the host, for a stationary and for a moving (ragged, NaN-padded) network."""
    N, k, n_cycles, S = 512, 12, 5, 20
    gen = dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True)
    nature = gen.generate(n_steps=n_cycles * S + 1)
    traj = np.ascontiguousarray(nature['x'].values)
    coords = {'time': nature['time'].values, 'index': np.arange(N)}
    kw = dict(times=nature['time'].values[np.arange(0, n_cycles * S + 1, S)], random_location_density=0.3, error_sd=0.8,
              random_seed=12, stationary_observers=stationary)
    obs_host = dab.observer.Observer(_xr.new_dataset({'x': (('time', 'index'), traj)}, coords=coords), **kw).observe()
    obs_dev = dab.observer.Observer(_xr.new_dataset({'x': (('time', 'index'), torch.from_numpy(traj).cuda())},
                                                    coords=coords), keep_on_device=True, **kw).observe()
    xv = _xr.var_data(obs_dev, 'x')
    assert xv.is_cuda and np.array_equal(xv.cpu().numpy(), obs_host['x'].values, equal_nan=True)
    if not stationary:
        assert np.isnan(obs_host['x'].values).any()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True))
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model, multiplicative_inflation=1.05)
    rng = np.random.default_rng(2)
    init = _xr.new_dataset({'x': (('ensemble', 'index'), traj[0] + rng.standard_normal((k, N)))}, coords={'time': 0.0})
    ckw = dict(input_state=init, start_time=0.0, n_cycles=n_cycles, obs_error_sd=0.8, analysis_window=0.2,
               analysis_time_in_window=0.0)
    a = dc.cycle(obs_vector=obs_host, **ckw)['x'].values
    b = dc.cycle(obs_vector=obs_dev, **ckw)['x'].values
    assert np.all(np.isfinite(a))
    if stationary:
        assert np.array_equal(a, b)
    else:
        # the host route drops NaN slots, the device route keeps them as weightless rows: same sums, other grouping
        assert rel_err(b, a) < 1e-10


# ------------------------------------------------------------------ ETKF.cycle() with the other generators
def _cycle_vs_oracle(dc, init_x, start_time, obs, sd, window, n_cycles, fc, delta_t, stationary=True,
                     return_forecast=False):
    init = _xr.new_dataset({'x': (('ensemble', 'index'), init_x)})
    out = dc.cycle(input_state=init, start_time=start_time, obs_vector=obs, obs_error_sd=sd,
                   analysis_window=window, n_cycles=n_cycles, return_forecast=return_forecast)
    ref = o_etkf.cycle(init_x, float(start_time), np.asarray(obs['x'].values).reshape(len(obs['time'].values), -1),
                       np.asarray(obs['time'].values), np.asarray(obs['system_index'].values)[0], n_cycles, sd, fc,
                       delta_t, stationary_observers=stationary, analysis_window=window,
                       return_forecast=return_forecast, rho=dc.multiplicative_inflation, literal=True)
    return np.asarray(out['x'].values), ref


def test_etkf_cycle_with_a_lorenz63_model():
    """The reference's ETKF takes any Model; here `Lorenz63Model`: analysis kernels and the
    Lorenz-63 ensemble forecast alternate on the device.  Against the oracle cycle with the same RK4."""
    from oracle import l63 as o_l63
    gen = dab.data.Lorenz63(delta_t=0.01)
    nature = gen.generate(n_steps=400)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(5, 400, 10)],
                                random_location_count=2, error_sd=0.7, random_seed=5,
                                stationary_observers=True).observe()
    model = dab.model.Lorenz63Model(model_obj=dab.data.Lorenz63(delta_t=0.01))
    dc = dab.dacycler.ETKF(system_dim=3, delta_t=0.01, ensemble_dim=6, model_obj=model, multiplicative_inflation=1.05)
    x0 = nature['x'].values[0] + threefry.normal(3, (6, 3))
    fc = lambda x, n: o_l63.generate_rk4(x, n, 0.01)
    out, ref = _cycle_vs_oracle(dc, x0, 0.0, obs, 0.7, 0.1, 30, fc, 0.01)
    assert out.shape == (30, 6, 3) and np.all(np.isfinite(out))
    assert rel_err(out, ref) < 1e-9
    # windows without observations (every second one): inflation only, not the empty-R collapse
    obs2 = dab.observer.Observer(nature, times=nature['time'].values[np.arange(5, 400, 20)],
                                 random_location_count=3, error_sd=0.7, random_seed=5).observe()
    out2, ref2 = _cycle_vs_oracle(dc, x0, 0.0, obs2, 0.7, 0.1, 10, fc, 0.01)
    assert rel_err(out2, ref2) < 1e-9
    assert np.std(out2[1], axis=0).min() > 1e-3
    # return_forecast=True: every cycle's trajectory without its last point
    out3, ref3 = _cycle_vs_oracle(dc, x0, 0.0, obs, 0.7, 0.1, 5, fc, 0.01, return_forecast=True)
    assert out3.shape == (5, 6, 10, 3) and rel_err(out3, ref3) < 1e-9


def test_etkf_cycle_with_an_sqgturb_model():
    """`SQGTurbModel` behind ETKF.cycle(): the flattened gridded potential vorticity is the state, every cycle is
    analysis -> rfft2 -> RK4 steps -> irfft2 on the device.  Against the oracle cycle whose forecast plugin is the
    oracle generator under the cyclers' trajectory contract (n points, the first is the analysis itself)."""
    from oracle import sqgturb as o_sqg
    from test_gpu_sqgturb import members
    N, k, sd = 16, 6, 50.0
    pv = members(N, k + 1, seed=9)
    gen = dab.data.SQGTurb(pv=pv[0], precision='double')
    nat = gen.generate(n_steps=12)
    flat = np.asarray(nat['pv'].values).reshape(13, -1)
    nature = _xr.new_dataset({'x': (('time', 'index'), flat)},
                             coords={'time': np.asarray(nat['time'].values), 'index': np.arange(flat.shape[1])},
                             attrs={'system_dim': flat.shape[1], 'delta_t': 900.0})
    obs = dab.observer.Observer(nature, times=np.asarray(nat['time'].values)[1::2], random_location_count=120,
                                error_sd=sd, random_seed=3).observe()
    model = dab.model.SQGTurbModel(model_obj=dab.data.SQGTurb(pv=pv[0], precision='double'))
    dc = dab.dacycler.ETKF(system_dim=2 * N * N, delta_t=900, ensemble_dim=k, model_obj=model)
    x0 = pv[1:].reshape(k, -1)
    T = o_sqg.tables(N, 'double')
    fc = lambda x, n: np.concatenate([x[None], o_sqg.integrate(T, np.fft.rfft2(x.reshape(2, N, N)), n - 1)[1]
                                      .reshape(n - 1, -1)])
    out, ref = _cycle_vs_oracle(dc, x0, 0.0, obs, sd, 1800.0, 4, fc, 900)
    assert out.shape == (4, k, 2 * N * N) and np.all(np.isfinite(out))
    assert rel_err(out, ref) < 1e-9
    out3, ref3 = _cycle_vs_oracle(dc, x0, 0.0, obs, sd, 1800.0, 2, fc, 900, return_forecast=True)
    assert out3.shape == (2, k, 2, 2 * N * N) and rel_err(out3, ref3) < 1e-9
    # the member-by-member plugin call the reference's ETKF would make gives the same forecast
    one = _xr.new_dataset({'x': (('index',), out[0, 0])})
    last, traj = model.forecast(one, n_steps=3)
    assert traj['x'].shape == (3, 2 * N * N) and np.array_equal(np.asarray(traj['x'].values)[0], out[0, 0])


def test_etkf_cycle_lorenz63_with_moving_observers():
    """The generic-model loop with a non-stationary network (ragged counts, NaN-padded slots masked,
    dabench/dacycler/_dacycler.py:262-268) against the oracle cycle."""
    from oracle import l63 as o_l63
    gen = dab.data.Lorenz63(delta_t=0.01)
    nature = gen.generate(n_steps=200)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(5, 200, 10)],
                                random_location_density=0.6, error_sd=0.5, random_seed=11,
                                stationary_observers=False).observe()
    vals = np.asarray(obs['x'].values)
    assert np.isnan(vals).any() and not np.isnan(vals).all()          # ragged: some slots are padding
    model = dab.model.Lorenz63Model(model_obj=dab.data.Lorenz63(delta_t=0.01))
    dc = dab.dacycler.ETKF(system_dim=3, delta_t=0.01, ensemble_dim=5, model_obj=model)
    x0 = nature['x'].values[0] + threefry.normal(8, (5, 3))
    init = _xr.new_dataset({'x': (('ensemble', 'index'), x0)})
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, obs_error_sd=0.5, analysis_window=0.1,
                   n_cycles=15)
    fc = lambda x, n: o_l63.generate_rk4(x, n, 0.01)
    ref = o_etkf.cycle(x0, 0.0, vals, np.asarray(obs['time'].values), np.asarray(obs['system_index'].values)[0], 15,
                       0.5, fc, 0.01, stationary_observers=False, analysis_window=0.1, literal=True)
    assert rel_err(np.asarray(out['x'].values), ref) < 1e-9


# ------------------------------------------------------------------ the reference's Lorenz-63 observer tests
def test_reference_observer_l63_recipes_through_the_package():
    """tests/observer_base_test.py:11-131 written against this package: the Lorenz-63 nature run comes from the GPU
    (fixed-step RK4 instead of the reference's adaptive Dormand-Prince: the value goldens hold to np.allclose), the
    seeded draws reproduce the reference's indices, times and errors."""
    l63 = dab.data.Lorenz63()
    ds = l63.generate(n_steps=10)
    x = np.asarray(ds['x'].values)
    # test_obs_l63
    ov = dab.observer.Observer(ds, random_time_density=0.5, random_location_density=0.5, error_sd=0.7,
                               random_seed=99).observe()
    assert ov['x'].values.shape[0] == 8 and ov['time'].shape[0] == 8
    assert np.array_equal(ov['system_index'].values, np.repeat(1, 8).reshape((1, 8, 1)))
    assert ov['x'].values[0, 0] == pytest.approx(x[0, 1] + np.asarray(ov['errors'].values)[0, 0, 0])
    assert np.allclose(ov['x'].values, np.array([[-14.77003751], [-14.82969835], [-16.49873315], [-16.21225914],
                                                 [-16.08467213], [-16.69890358], [-15.45879053], [-15.12257015]]))
    assert np.allclose(ov['errors'].values, np.array([[[0.22996249], [0.65467059], [-0.6143291], [-0.03212726],
                                                       [0.26731022], [-0.31677728], [0.50516542], [-0.24651439]]]))
    assert np.array_equal(ov['time'].values, np.array([0., 0.01, 0.02, 0.03, 0.04, 0.05, 0.07, 0.09]))
    # test_obs_l63_count
    ov = dab.observer.Observer(ds, random_time_count=5, random_location_count=2, error_sd=0.7).observe()
    assert ov['x'].shape[0] == 5 and ov['time'].shape[0] == 5
    assert np.array_equal(ov['system_index'].values, np.tile([1, 2], 5).reshape((1, 5, 2)))
    assert np.allclose(ov['x'].values, np.array([[-16.71412359, 23.46140116], [-16.5006219, 24.10746168],
                                                 [-17.11500315, 27.69793923], [-15.78352562, 29.22759993],
                                                 [-14.87744996, 31.11579368]]))
    assert np.allclose(np.asarray(ov['errors'].values)[0], np.array([[-1.22975335, 1.17910212],
                                                                     [-0.32048999, -0.41749407],
                                                                     [-0.73287729, 0.65225442],
                                                                     [0.47248634, 0.87110884],
                                                                     [0.6251612, 0.18410346]]))
    assert np.array_equal(ov['time'].values, np.array([0.01, 0.03, 0.05, 0.06, 0.08]))
    # test_obs_l63_diffseed
    ov = dab.observer.Observer(ds, random_time_density=0.5, random_location_density=0.5, error_sd=0.0,
                               random_seed=1).observe()
    assert np.array_equal(ov['system_index'].values, np.tile([0, 2], 5).reshape((1, 5, 2)))
    assert ov['x'].values[0, 0] == x[0, 0]
    assert np.array_equal(ov['time'].values, np.array([0., 0.01, 0.03, 0.06, 0.08]))
    assert np.allclose(ov['x'].values, x[[0, 1, 3, 6, 8]][:, [0, 2]])
    assert np.array_equal(ov['errors'].values, np.repeat(0, 10).reshape(1, 5, 2))
    # test_obs_l63_specific_locs
    ov = dab.observer.Observer(ds, locations={'index': np.array([2])}, times=ds['time'].values[[4, 9]],
                               error_sd=1.0).observe()
    assert np.array_equal(ov['system_index'].values, np.repeat(2, 2).reshape((1, 2, 1)))
    assert np.allclose(np.asarray(ov['x'].values) - np.asarray(ov['errors'].values)[0], np.array([[x[4, 2]], [x[9, 2]]]))
    assert np.array_equal(ov['time'].values, np.array([0.04, 0.09]))
    assert np.allclose(np.asarray(ov['errors'].values)[0], np.array([[0.0824943], [-0.46441841]]))


def test_etkf_cycle_lorenz63_with_a_sigma_per_observation():
    """A vector obs_error_sd (diagonal R with unequal entries, one weightless row) through the generic-model loop,
    against the oracle analysis with the same per-observation weights and the oracle Lorenz-63 RK4."""
    from oracle import l63 as o_l63
    gen = dab.data.Lorenz63(delta_t=0.01)
    nature = gen.generate(n_steps=100)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, 100, 10)],
                                random_location_count=3, error_sd=0.5, random_seed=2).observe()
    sdv = np.array([0.4, 0.0, 1.3])
    model = dab.model.Lorenz63Model(model_obj=dab.data.Lorenz63(delta_t=0.01))
    dc = dab.dacycler.ETKF(system_dim=3, delta_t=0.01, ensemble_dim=7, model_obj=model)
    X0 = nature['x'].values[0] + threefry.normal(21, (7, 3))
    init = _xr.new_dataset({'x': (('ensemble', 'index'), X0)})
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=8, obs_error_sd=sdv,
                   analysis_window=0.1, analysis_time_in_window=0.0)
    w = np.where(sdv == 0.0, 0.0, 1.0 / np.where(sdv == 0.0, 1.0, sdv) ** 2)
    loc = np.asarray(obs['system_index'].values)[0][0]
    Xc = X0.copy()
    for c in range(8):
        Xa = o_etkf.analysis_sparse(Xc, np.asarray(obs['x'].values)[c], loc, w, 1.0)
        assert rel_err(np.asarray(out['x'].values)[c], Xa) < 1e-9, c
        Xc = o_l63.rk4_forecast(Xa, 10, 0.01)


def test_var3d_cycle_with_lorenz63_and_sqgturb_models():
    """Var3D behind DACycler.cycle with the other generators (the reference's Var3D takes any Model): analysis on
    the device, the model's own device forecast with one member; against the oracle's literal Var3D cycle with the
    oracle generator as the plugin."""
    from oracle import l63 as o_l63, sqgturb as o_sqg
    from test_gpu_sqgturb import members
    # Lorenz-63
    nature = dab.data.Lorenz63(delta_t=0.01).generate(n_steps=120)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, 120, 10)], random_location_count=2,
                                error_sd=0.5, random_seed=4).observe()
    dc = dab.dacycler.Var3D(system_dim=3, delta_t=0.01, model_obj=dab.model.Lorenz63Model(
        model_obj=dab.data.Lorenz63(delta_t=0.01)))
    x0 = nature['x'].values[0] + np.array([0.5, -0.3, 0.8])
    init = _xr.new_dataset({'x': (('index',), x0.copy())})
    full = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=10, obs_error_sd=0.5,
                    analysis_window=0.1, analysis_time_in_window=0.0, return_forecast=True)
    fc = lambda x, n: o_l63.generate_rk4(x, n, 0.01)
    ref = o_var3d.cycle(x0, 0.0, obs['x'].values, obs['time'].values, np.asarray(obs['system_index'].values)[0], 10,
                        0.5, fc, 0.01, analysis_window=0.1, analysis_time_in_window=0.0, return_forecast=True,
                        literal=True)
    assert full['x'].shape == (10, 10, 3) and rel_err(np.asarray(full['x'].values), ref) < 1e-9
    # SQGTurb, flattened gridded state
    N = 16
    pv = members(N, 2, seed=13)
    gen = dab.data.SQGTurb(pv=pv[0], precision='double')
    nat = gen.generate(n_steps=8)
    flat = np.asarray(nat['pv'].values).reshape(9, -1)
    nds = _xr.new_dataset({'x': (('time', 'index'), flat)},
                          coords={'time': np.asarray(nat['time'].values), 'index': np.arange(flat.shape[1])})
    obs = dab.observer.Observer(nds, times=np.asarray(nat['time'].values)[0::2][:4], random_location_count=60,
                                error_sd=40.0, random_seed=6).observe()
    dc = dab.dacycler.Var3D(system_dim=2 * N * N, delta_t=900, model_obj=dab.model.SQGTurbModel(
        model_obj=dab.data.SQGTurb(pv=pv[0], precision='double')))
    x0 = pv[1].ravel()
    init = _xr.new_dataset({'x': (('index',), x0.copy())})
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=3, obs_error_sd=40.0,
                   analysis_window=1800.0, analysis_time_in_window=0.0)
    T = o_sqg.tables(N, 'double')
    fc = lambda x, n: np.concatenate([x[None], o_sqg.integrate(T, np.fft.rfft2(x.reshape(2, N, N)), n - 1)[1]
                                      .reshape(n - 1, -1)])
    ref = o_var3d.cycle(x0, 0.0, obs['x'].values, obs['time'].values, np.asarray(obs['system_index'].values)[0], 3,
                        40.0, fc, 900, analysis_window=1800.0, analysis_time_in_window=0.0, literal=False)
    assert out['x'].shape == (3, 2 * N * N) and rel_err(np.asarray(out['x'].values), ref) < 1e-9
