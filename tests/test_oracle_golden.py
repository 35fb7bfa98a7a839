"""Pins the oracle against every known-answer vector the reference's own tests hold for the
ETKF hot path.  The literal numbers below are the reference's test
data, cited by file:line; nothing here needs a GPU."""
import numpy as np
import pytest
import scipy.integrate

from oracle import l96, l63, etkf, var3d, observer, threefry, metrics, windows


# ---------------------------------------------------------------- ETKF end-to-end golden
@pytest.fixture(scope="module")
def etkf_golden_run():
    """Recipe of /root/reference/tests/dacycler_etkf_test.py:13-75."""
    x0 = l96.default_x0(5, 0.01)
    nature = l96.generate_dopri5(x0, 25, 0.01)                       # :16-17
    t = l96.time_grid(25, 0.01)
    obs = observer.observe(nature, t, times=t[np.arange(0, 25, 5)],  # :22-33
                           random_location_count=3, error_bias=0.1, error_sd=1.0,
                           random_seed=91, stationary_observers=True)
    noise = threefry.normal(42, (8, 5))                              # :10,60
    init = nature[10] + noise
    fc = lambda x, n: l96.generate_dopri5(x, n, 0.05)                # model delta_t=0.05, :37
    run = lambda literal: etkf.cycle(
        init, t[10], obs['values'], obs['time'], obs['system_index'],
        n_cycles=10, obs_error_sd=1.5, forecast_fn=fc, delta_t=0.01,
        analysis_window=0.1, return_forecast=True, literal=literal)
    return obs, run(True), run(False)


def test_etkf_reference_golden(etkf_golden_run):
    obs, out, _ = etkf_golden_run
    assert list(obs['system_index'][0]) == [2, 1, 4]
    assert out.shape == (10, 8, 10, 5)
    flat = out.transpose(0, 2, 1, 3).reshape(100, 8, 5)              # stack(time=[cycle, step])
    assert not np.allclose(flat[-1, 1], flat[-1, 0])                 # :84-87
    # /root/reference/tests/dacycler_etkf_test.py:89-92
    assert np.allclose(flat[0, 0], [-0.85402591, 1.03480315, 0.51005132, 6.61546551, 8.1166806])
    # :94-97
    assert np.allclose(flat[-1, 0], [0.66697948, 3.15465627, 5.39288975, -4.96130847, 2.17202611])
    # :99-102
    assert np.allclose(flat[-1].mean(axis=0),
                       [1.45024252, 3.81627191, 5.4507981, 1.21646539, 0.09439264])


def test_sparse_formulation_equals_literal_in_cycling(etkf_golden_run):
    _, lit, sparse = etkf_golden_run
    assert np.max(np.abs(lit - sparse)) / np.max(np.abs(lit)) < 1e-10


@pytest.mark.parametrize("N,k,n_y,rho", [(36, 10, 18, 1.0), (40, 20, 20, 1.1),
                                         (2000, 64, 600, 1.0), (1500, 128, 1500, 1.05)])
def test_sparse_formulation_equals_literal_single_analysis(N, k, n_y, rho):
    rng = np.random.default_rng(N + k)
    truth = 3.0 * rng.standard_normal(N) + 2.0
    X = truth + rng.standard_normal((k, N))
    idx = rng.choice(N, n_y, replace=False, shuffle=False)
    y = truth[idx] + rng.standard_normal(n_y)
    mask = rng.random(n_y) > 0.1
    sd = 0.8
    H = etkf.calc_default_H(n_y, N, idx)
    H = np.where(mask, H.T, 0).T
    lit = etkf.analysis_literal(X.T, y, H, etkf.calc_default_R(n_y, sd), rho).T
    sp = etkf.analysis_sparse(X, y, idx, mask / sd ** 2, rho)
    assert np.max(np.abs(lit - sp)) / np.max(np.abs(lit)) < 1e-12


def test_no_obs_cycle_is_sqrt_rho_inflation():
    """with all rows masked Wa = sqrt(rho) I, wa = 0."""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 12))
    Xa = etkf.analysis_sparse(X, np.zeros(4), np.zeros(4, dtype=int), np.zeros(4), rho=1.21)
    m = X.mean(axis=0)
    assert np.allclose(Xa, m + 1.1 * (X - m), atol=1e-13)
    # empty R (no observation slots at all): the reference's zero branch collapses to the mean
    Xe = etkf.analysis_literal(X.T, np.zeros(0), np.zeros((0, 12)), np.zeros((0, 0)), 1.0).T
    assert np.allclose(Xe, np.tile(m, (6, 1)))
    assert np.allclose(etkf.analysis_sparse(X, [], [], [], empty_R=True), Xe)


# ---------------------------------------------------------------- Lorenz-96 golden
# /root/reference/tests/data_lorenz96_test.py:120-131
L96_TRUE = np.array(
    [[8., 8., 8.01, 8., 8.],
     [-0.98989847, 3.37430623, 11.26913436, 12.84401465, -0.72462902],
     [2.61054548, -0.07543557, 0.32725793, 11.42165675, 1.87837841],
     [3.20376924, 6.83403523, -2.63507472, 3.05415108, 2.56797523],
     [0.20204786, 1.26443727, 8.39982716, -2.45623576, 0.39051976],
     [-5.27169559, 0.97322104, 2.41156636, 6.96395228, -1.70612383],
     [4.26726853, 6.23099434, -1.22264959, 2.29152838, 2.84244614],
     [-0.25361909, 0.92942401, 8.63986765, -1.45411412, -0.89528154],
     [-5.27402622, -1.10262156, -0.53270967, 7.58069989, 2.84768712],
     [4.73878926, 6.27758562, -2.10858776, 3.26147038, 3.168998]])


def test_l96_reference_trajectory():
    """The reference default store_as_jax=False integrates with scipy odeint (LSODA),
    dabench/data/_utils.py:53; t_final=10, dt=1e-3."""
    x0 = 8.0 * np.ones(5)
    x0[2] += 0.01
    t = np.arange(0.0, 10 - 0.001 / 2, 0.001)
    y = scipy.integrate.odeint(lambda x, _t: l96.rhs(x), x0, t)[::1000]
    assert y.shape == L96_TRUE.shape
    assert np.allclose(y, L96_TRUE, rtol=0.01, atol=0)               # :135
    assert np.allclose(np.sum(y), np.sum(L96_TRUE), rtol=1e-3, atol=0)


def test_l63_reference_known_answer():
    """/root/reference/tests/model_base_test.py:18-30: Lorenz63().generate(x0, n_steps=2)[-1]."""
    gold = np.array([-10.499835, -15.48437, 22.282299])
    assert np.allclose(l63.X0_DEFAULT, np.array([-10., -15., 21.3]))
    assert np.allclose(l63.generate_dopri5(l63.X0_DEFAULT, 2, 0.01)[-1], gold)
    assert np.allclose(l63.generate_rk4(l63.X0_DEFAULT, 2, 0.01)[-1], gold)          # the kernel's integrator
    # the restated adaptive integrator against an independent one
    ref = scipy.integrate.solve_ivp(lambda t, y: l63.rhs(y), (0.0, 0.5), l63.X0_DEFAULT, rtol=1e-12,
                                    atol=1e-12, t_eval=np.arange(50) * 0.01).y.T
    assert np.max(np.abs(l63.generate_dopri5(l63.X0_DEFAULT, 50, 0.01) - ref)) < 5e-6
    # an ensemble is advanced member-wise
    X = l63.X0_DEFAULT + np.arange(12).reshape(4, 3) * 0.1
    assert np.array_equal(l63.rk4_forecast(X, 5, 0.01)[2], l63.rk4_forecast(X[2], 5, 0.01))


def test_rhs_matches_explicit_index_form():
    """dabench/data/_lorenz96.py:138-149 written with explicit indices."""
    rng = np.random.default_rng(1)
    for N in (4, 5, 36, 40):
        x = rng.standard_normal(N)
        ref = np.empty(N)
        for i in range(N):
            ref[i] = (x[(i + 1) % N] - x[(i - 2) % N]) * x[(i - 1) % N] - x[i]
        ref += 8.0
        assert np.array_equal(l96.rhs(x), ref)


def test_dopri5_agrees_with_scipy_rk45():
    x0 = l96.default_x0(36, 0.01)
    a = l96.generate_dopri5(x0, 21, 0.01)
    t = l96.time_grid(21, 0.01)
    b = scipy.integrate.solve_ivp(lambda _t, x: l96.rhs(x), (0, t[-1]), x0, t_eval=t,
                                  method='RK45', rtol=1e-11, atol=1e-11).y.T
    assert np.max(np.abs(a - b)) < 1e-6
    c = l96.generate_rk4(x0, 21, 0.01)
    assert a.shape == c.shape == (21, 36)
    assert np.max(np.abs(c - b)) < 1e-4                               # RK4 @ dt=0.01: ~2.5e-6


# ---------------------------------------------------------------- observer goldens
def _l96_default_run():
    # data.Lorenz96() defaults: N=36, dt=0.05, store_as_jax=False -> scipy odeint
    x0 = l96.default_x0(36, 0.05)
    t = l96.time_grid(10, 0.05)
    return scipy.integrate.odeint(lambda x, _t: l96.rhs(x), x0, t), t


def test_observer_stationary_golden():
    """/root/reference/tests/observer_base_test.py:133-157"""
    traj, t = _l96_default_run()
    o = observer.observe(traj, t, random_time_density=0.4, random_location_density=0.2,
                         error_sd=0.7)
    assert o['values'].shape == (3, 4)
    assert np.array_equal(o['system_index'], np.tile([16, 9, 31, 32], 3).reshape(3, 4))
    assert np.allclose(o['time'], [0.15, 0.2, 0.45])
    assert np.allclose(o['errors'][0], [0.5913821, -0.22663247, 0.00787655, -0.29022197])
    assert np.allclose(o['values'][0], [6.09641725, 2.80249778, 2.8026745, 6.91684992])
    assert o['values'][0, 0] == pytest.approx(traj[3, 16] + o['errors'][0, 0])


def test_observer_nonstationary_golden():
    """/root/reference/tests/observer_base_test.py:160-180"""
    traj, t = _l96_default_run()
    o = observer.observe(traj, t, random_time_density=0.4, random_location_density=0.3,
                         error_sd=1.2, stationary_observers=False)
    assert o['values'].shape == (3, 12)
    assert o['system_index'][2, -1] == 32
    assert o['system_index'][0, 0] == 10
    assert o['values'][0, 0] == pytest.approx(traj[3, 10] + o['errors'][0, 0])
    assert o['values'][2, 5] == pytest.approx(0.010236507277464613)
    assert o['errors'][1, 4] == pytest.approx(-0.7470301078422978)
    assert np.allclose(o['time'], [0.15, 0.2, 0.45])
    assert np.isnan(o['values'][0, 8:]).all() and (o['system_index'][0, 8:] == 0).all()


# ---------------------------------------------------------------- windows + metrics
def test_window_bookkeeping():
    obs_t = np.round(np.arange(0, 1.0, 0.05), 10)
    times = windows.get_all_times(0.1, 0.1, 4)
    assert np.allclose(times, [0.1, 0.2, 0.3, 0.4])
    idx = windows.get_obs_indices(times, obs_t, 0.1)
    # [t-0.05, t+0.05): start inclusive, end exclusive
    assert [list(i) for i in idx] == [[1, 2], [3, 4], [5, 6], [7, 8]]
    pad = windows.pad_time_indices(idx + [np.array([], dtype=int)])
    assert pad.shape == (5, 2) and list(pad[0]) == [2, 3] and list(pad[-1]) == [0, 0]


def test_metrics():
    a = np.array([1.0, 2.0, 4.0])
    b = np.array([1.5, 2.0, 2.0])
    assert metrics.mse(a, b) == pytest.approx((0.25 + 4) / 3)
    assert metrics.rmse(a, b) == pytest.approx(np.sqrt((0.25 + 4) / 3))
    assert metrics.mae(a, b) == pytest.approx(2.5 / 3)


def test_get_all_times_keeps_reference_float_quirk():
    """dabench/dacycler/_utils.py:32-36 builds the window centres with
    arange(0, n*w, w); for e.g. n=12, w=0.1 that arange has 13 elements and the reference
    fails on the broadcast.  The restatement keeps that behaviour on purpose."""
    assert len(windows.get_all_times(0.0, 0.1, 10)) == 10
    with pytest.raises(ValueError):
        windows.get_all_times(0.0, 0.1, 12)


def test_var3d_reference_golden():
    """/root/reference/tests/dacycler_var3d_test.py:11-84 with the oracle: Lorenz-96 N=6, dt=0.05, 34 of 50
    observation times (density 0.7, seed 94), 3 stationary locations, sd 0.7, 10 cycles of window 0.25.
    The literal analysis (dense H/R/B, conjugate gradients as jax.scipy.sparse.linalg.cg) and the closed
    form for the default operators both reproduce the two presaved vectors to the printed digits."""
    x0 = l96.default_x0(6, 0.05)
    nature = l96.generate_dopri5(x0, 50, 0.05)
    times = l96.time_grid(50, 0.05)
    obs = observer.observe(nature, times, random_time_density=0.7, random_location_count=3, error_bias=0.0,
                           error_sd=0.7, random_seed=94, stationary_observers=True)
    assert obs['values'].shape == (34, 3)
    init = nature[0] + threefry.normal(42, (6,))
    fc = lambda x, n: l96.generate_dopri5(x, n, 0.05)
    first = np.array([-0.90632236, -1.20861455, 1.64865068, 5.11034063, 4.399881, -3.75779771])
    last = np.array([3.92060079, 3.97290102, -0.763032, -1.5979558, -0.0086728, 2.60395146])
    runs = []
    for literal in (True, False):
        out = var3d.cycle(init, times[0], obs['values'], obs['time'], obs['system_index'], 10, 0.7, fc, 0.05,
                          analysis_window=0.25, literal=literal)
        assert out.shape == (10, 6)
        assert np.max(np.abs(out[0] - first)) < 2e-8 and np.max(np.abs(out[-1] - last)) < 1e-7
        runs.append(out)
    assert np.max(np.abs(runs[0] - runs[1])) < 1e-9          # CG on a diagonal system is exact in <= 3 steps


def test_var3d_cg_matches_direct_solve():
    rng = np.random.default_rng(3)
    n, m = 40, 25
    xb, y = rng.standard_normal(n), rng.standard_normal(m)
    loc = rng.integers(0, n, m)                               # repeated locations on purpose
    H = etkf.calc_default_H(m, n, loc[None])
    H[::7] = 0.0                                              # masked rows
    R = etkf.calc_default_R(m, 0.8)
    xa = var3d.analysis_literal(xb, y, H, R)
    A = np.identity(n) + H.T @ H / 0.64
    assert np.max(np.abs(xa - np.linalg.solve(A, xb + H.T @ y / 0.64))) < 1e-4     # CG tolerance 1e-5
    w = np.ones(m) / 0.64
    w[::7] = 0.0
    assert np.max(np.abs(var3d.analysis_diagonal(xb, y, loc, w) - np.linalg.solve(A, xb + H.T @ y / 0.64))) < 1e-13
