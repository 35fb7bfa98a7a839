"""Batched ETKF experiments (BASELINE.json configs[1]): many independent small `ETKF.cycle()` runs --
e.g. an inflation sweep -- in ONE kernel launch, one CTA per experiment
(`dab_etkf_cycle_batched_f64`).  The reference has no batched entry point (it would loop over
`cycle()` calls); this helper takes the same objects `cycle()` takes, one per experiment.
"""
import numpy as np

from .. import _xr


def cycle_batched(cyclers, input_states, start_time, obs_vectors, n_cycles, obs_error_sd=None,
                  analysis_window=0.2, analysis_time_in_window=None, return_mean_only=False):
    """Run `len(cyclers)` independent experiments.

    cyclers      : list of `ETKF` objects of identical shape (they may differ in
                   `multiplicative_inflation`)
    input_states : list of Datasets `x[ensemble, index]` (or one, shared)
    obs_vectors  : list of observation Datasets of identical shape and times (or one, shared)
    Returns a Dataset with `x[experiment, cycle, ensemble, index]` (analyses), or with
    return_mean_only=True `x_mean[experiment, cycle, index]` (ensemble-mean analyses).
    """
    from .._runtime import handle
    B = len(cyclers)
    states = list(input_states) if isinstance(input_states, (list, tuple)) else [input_states] * B
    obs = list(obs_vectors) if isinstance(obs_vectors, (list, tuple)) else [obs_vectors] * B
    shared_obs = not isinstance(obs_vectors, (list, tuple))
    plans = [c._plan(s, start_time, o, n_cycles, obs_error_sd, analysis_window, analysis_time_in_window)
             for c, s, o in zip(cyclers, states, obs)]
    cfg = plans[0]['cfg']
    for p in plans[1:]:
        q = p['cfg']
        same = all(getattr(cfg, f) == getattr(q, f) for f, _ in cfg._fields_ if f != 'rho')
        if not same or not np.array_equal(p['padded'], plans[0]['padded']):
            raise ValueError('all experiments of a batch must have the same shape, windows and model')
    k, N = cfg.ensemble_dim, cfg.system_dim
    rho = np.ascontiguousarray([p['cfg'].rho for p in plans], dtype=np.float64)
    x0 = np.ascontiguousarray(np.stack([p['X0'] for p in plans]))
    if shared_obs:
        vals = plans[0]['vals']                    # one observation set, uploaded once (flag bits 0 and 1)
        idx = plans[0]['sysidx']
    else:
        vals = np.ascontiguousarray(np.stack([p['vals'] for p in plans]))
        idx = np.ascontiguousarray(np.stack([p['sysidx'] for p in plans]))
    H = handle()
    dim = plans[0]['state_dim']
    if return_mean_only:
        mean = np.empty((B, n_cycles, N))
        H.etkf_cycle_batched(cfg, B, rho, x0, vals, idx, 3 if shared_obs else 0, plans[0]['padded'], None, mean, None)
        return _xr.new_dataset({'x_mean': (('experiment', 'cycle', dim), mean)})
    out = np.empty((B, n_cycles, k, N))
    H.etkf_cycle_batched(cfg, B, rho, x0, vals, idx, 3 if shared_obs else 0, plans[0]['padded'], out, None, None)
    return _xr.new_dataset({'x': (('experiment', 'cycle', 'ensemble', dim), out)})
