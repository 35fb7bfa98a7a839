"""Generates the committed golden fixtures from the CPU oracle (run here, once; the oracle's
reference-faithful Dormand-Prince integrator is too slow to run inside the GPU tests).

  python tests/golden/make_golden.py

etkf_converged_dp5.npz -- the converged-filter experiment of SURVEY.md 7.3/8d (P4): Lorenz-96
  N=40, k=20, window 0.05 (S=5 steps of dt=0.01), all 40 points observed (stationary, random
  order), sigma=1, rho=1.05, 200 cycles; cycled with the oracle's LITERAL analysis and the
  reference-style adaptive Dormand-Prince forecast.  Stores inputs and the per-cycle analysis
  RMSE of the ensemble mean against the truth.
analysis_cases.npz -- single analyses computed with the oracle's literal formula
  (dense H, dense R, pinv, sqrtm) at C1 / C2 shapes and two down-scaled C3 / C5 shapes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import l96, etkf, metrics  # noqa: E402


def converged_experiment():
    N, k, S, dt, n_cycles, sd, rho = 40, 20, 5, 0.01, 200, 1.0, 1.05
    rng = np.random.default_rng(2024)
    x = l96.rk4_forecast(8.0 + rng.standard_normal(N), 2000, dt)            # spin-up
    truth = l96.generate_dopri5(x, n_cycles * S + 1, dt)                    # [n*S+1, N]
    t = l96.time_grid(n_cycles * S + 1, dt)
    tidx = np.arange(0, n_cycles * S + 1, S)
    locs = rng.choice(N, N, replace=False, shuffle=False)
    vals = truth[tidx][:, locs] + sd * rng.standard_normal((len(tidx), N))
    sysidx = np.tile(locs, (len(tidx), 1)).astype(np.int64)
    X0 = truth[0] + rng.standard_normal((k, N))
    # window centres are built as the product does (start + c*w); the reference's arange(0, n*w, w)
    # has 201 elements for n=200, w=0.05 and would fail on its own broadcast (windows.get_all_times)
    fc = lambda xx, n: l96.generate_dopri5(xx, n, dt)
    out = np.empty((n_cycles, k, N))
    X = X0
    w = np.full(N, 1.0 / sd ** 2)
    for c in range(n_cycles):
        H = etkf.calc_default_H(N, N, sysidx[c])
        Xa = etkf.analysis_literal(X.T, vals[c], H, etkf.calc_default_R(N, sd), rho).T
        out[c] = Xa
        X = np.stack([fc(Xa[i], S + 1)[-1] for i in range(k)])
    rmse = np.array([metrics.rmse(out[c].mean(axis=0), truth[tidx[c]]) for c in range(n_cycles)])
    print('converged experiment: time-mean analysis RMSE (cycles 50..199) = %.4f' % rmse[50:].mean())
    np.savez_compressed(os.path.join(HERE, 'etkf_converged_dp5.npz'), X0=X0, truth=truth[tidx[:n_cycles]],
                        vals=vals, sysidx=sysidx, obs_times=t[tidx], rmse_dp5=rmse,
                        params=np.array([N, k, S, dt, n_cycles, sd, rho]))


def analysis_cases():
    out = {}
    for tag, (N, k, n_y, rho, sd) in {'c1': (36, 10, 18, 1.0, 1.0), 'c2': (40, 20, 20, 1.3, 1.0),
                                      'c3s': (400, 64, 150, 1.0, 1.5), 'c5s': (200, 128, 200, 1.05, 0.8)}.items():
        rng = np.random.default_rng(n_y)
        truth = l96.rk4_forecast(8.0 + rng.standard_normal(N), 300, 0.01)
        X = truth + rng.standard_normal((k, N))
        idx = rng.choice(N, n_y, replace=False, shuffle=False).astype(np.int64)
        y = truth[idx] + sd * rng.standard_normal(n_y)
        mask = (rng.random(n_y) > 0.15).astype(np.uint8)
        H = etkf.calc_default_H(n_y, N, idx)
        H = np.where(mask, H.T, 0).T
        Xa = etkf.analysis_literal(X.T, y, H, etkf.calc_default_R(n_y, sd), rho).T
        for name, arr in dict(X=X, idx=idx, y=y, mask=mask, Xa=Xa, par=np.array([rho, sd])).items():
            out['%s_%s' % (tag, name)] = arr
    np.savez_compressed(os.path.join(HERE, 'analysis_cases.npz'), **out)


if __name__ == '__main__':
    analysis_cases()
    converged_experiment()
