"""Batched small-problem path (BASELINE.json configs[1]): every experiment of the batch must equal
the oracle's cycling of that experiment alone (RK4 forecast, sparse analysis)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dataassimbench_b200 import _native
from oracle import l96 as o_l96, etkf as o_etkf, windows as o_win
from test_gpu_kernels import cycler_case, rel_err


def run_batch(H, cases, rhos, N, k, n_cycles, S, dt, sd, stationary, shared_index):
    c0 = cases[0]
    times = o_win.get_all_times(c0['start'], c0['window'], n_cycles)
    padded = np.ascontiguousarray(o_win.pad_time_indices(o_win.get_obs_indices(times, c0['obs_times'], c0['window'])))
    width = max(c['vals'].shape[1] for c in cases)
    nt = c0['vals'].shape[0]
    B = len(cases)
    vals = np.full((B, nt, width), np.nan)
    idx = np.zeros((B, nt, width), dtype=np.int64)
    for b, c in enumerate(cases):
        vals[b, :, :c['vals'].shape[1]] = c['vals']
        idx[b, :, :c['vals'].shape[1]] = c['sysidx']
    x0 = np.ascontiguousarray(np.stack([c['X0'] for c in cases]))
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=1.0,
                              obs_error_sd=sd, n_obs_times=nt, n_obs=width, stationary=int(stationary),
                              n_cycles=n_cycles, t_pad=padded.shape[1], rank=0, world=1)
    out = np.empty((B, n_cycles, k, N))
    mean = np.empty((B, n_cycles, N))
    fin = np.empty((B, k, N))
    # shared_index: bit 0 = one index set, bit 1 = one set of observed values for all experiments
    H.etkf_cycle_batched(cfg, B, np.ascontiguousarray(rhos, dtype=np.float64), x0,
                         np.ascontiguousarray(vals[0]) if int(shared_index) & 2 else vals,
                         np.ascontiguousarray(idx[0]) if int(shared_index) & 1 else idx, int(shared_index), padded,
                         out, mean, fin)
    return out, mean, fin


@pytest.mark.parametrize("N,k,n_obs,n_cycles,S,stationary,obs_every", [
    (40, 20, 20, 10, 20, True, None),       # the C2 shape
    (40, 20, 20, 10, 20, False, None),      # non-stationary: ragged rows, NaN padding
    (36, 10, 18, 10, 20, True, None),       # the C1 shape, as a batch
    (40, 20, 40, 20, 5, True, None),        # converged regime
    (64, 32, 24, 5, 10, True, 5),           # k = 32 (largest), two obs times per window
    (5, 8, 3, 5, 10, True, None),           # the reference test's tiny shape
    (100, 20, 30, 3, 8, True, None),        # 16 consecutive points per thread (7 threads per member, ragged)
    (132, 31, 12, 2, 5, True, None),        # k * ceil(N/16) > 256: the round-robin point mapping
])
def test_batched_equals_oracle_per_experiment(N, k, n_obs, n_cycles, S, stationary, obs_every):
    sd, dt = 1.0, 0.01
    rhos = [1.0, 1.05, 1.3]
    cases = [cycler_case(N, k, n_obs, n_cycles, S, sd, r, stationary, seed=100 + i, obs_every=obs_every)
             for i, r in enumerate(rhos)]
    H = _native.Handle(0)
    out, mean, fin = run_batch(H, cases, rhos, N, k, n_cycles, S, dt, sd, stationary, shared_index=False)
    H.close()
    fc = lambda x, n: o_l96.generate_rk4(x, n, dt)
    for b, (c, rho) in enumerate(zip(cases, rhos)):
        ref = o_etkf.cycle(c['X0'], c['start'], c['vals'], c['obs_times'], c['sysidx'], n_cycles, sd, fc, dt,
                           stationary_observers=stationary, analysis_window=c['window'],
                           return_forecast=True, rho=rho, literal=False)
        assert rel_err(out[b], ref[:, :, 0, :]) < 1e-9
        assert rel_err(mean[b], ref[:, :, 0, :].mean(axis=1)) < 1e-9
    assert np.all(np.isfinite(fin))


def test_batched_inflation_sweep_shared_index_matches_unbatched_gpu():
    """Same experiment with 8 inflation factors (shared observation network): each batch member
    equals the large-path cycler run with that rho."""
    from test_gpu_kernels import run_gpu_cycler
    N, k, n_obs, n_cycles, S, sd = 40, 20, 20, 20, 20, 1.0
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, 1.0, True, seed=9)
    rhos = np.linspace(1.0, 1.5, 8)
    H = _native.Handle(0)
    out, _, fin = run_batch(H, [case] * len(rhos), rhos, N, k, n_cycles, S, 0.01, sd, True, shared_index=True)
    for b, rho in enumerate(rhos):
        ref, ref_fin, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, float(rho), True)
        assert rel_err(out[b], ref) < 1e-10
        assert rel_err(fin[b], ref_fin) < 1e-9
    # one upload of the observed values for the whole sweep (flag bit 1): the same bits
    out3, _, fin3 = run_batch(H, [case] * len(rhos), rhos, N, k, n_cycles, S, 0.01, sd, True, shared_index=3)
    assert np.array_equal(out3, out) and np.array_equal(fin3, fin)
    H.close()


def test_batched_argument_errors():
    H = _native.Handle(0)
    cfg = _native.CycleConfig(system_dim=40, ensemble_dim=64, n_steps=5, model_dt=0.01, forcing=8.0, rho=1.0,
                              obs_error_sd=1.0, n_obs_times=1, n_obs=4, stationary=1, n_cycles=1, t_pad=1,
                              rank=0, world=1)
    with pytest.raises(_native.DabError):      # k > 32 is the large path's job
        H.etkf_cycle_batched(cfg, 1, np.ones(1), np.zeros((1, 64, 40)), np.zeros((1, 1, 4)),
                             np.zeros((1, 1, 4), dtype=np.int64), False, np.ones((1, 1), dtype=np.int64))
    H.close()
