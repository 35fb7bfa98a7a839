"""C2 timing: 4096 independent Lorenz-96 N=40, k=20 ETKF experiments (inflation sweep), 200 cycles,
S=20, one launch.  Prints experiments-cycles/s and member-steps/s, plus the oracle on a sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dataassimbench_b200 import _native
from oracle import l96 as o_l96, etkf as o_etkf

B = int(os.environ.get("B", 4096)); N, k, S, n_cycles, n_obs, sd, dt = 40, 20, 20, 200, 20, 1.0, 0.01
rng = np.random.default_rng(1)
truth0 = o_l96.rk4_forecast(8 + rng.standard_normal(N), 1000, dt)
_, truth = o_l96.rk4_forecast(truth0, n_cycles * S, dt, return_traj=True)
tidx = np.arange(0, n_cycles * S + 1, S)
locs = rng.choice(N, n_obs, replace=False, shuffle=False)
vals1 = truth[tidx][:, locs] + sd * rng.standard_normal((len(tidx), n_obs))
vals1 = np.ascontiguousarray(vals1)          # one observation set for the whole sweep (shared_index = 3)
idx = np.ascontiguousarray(np.tile(locs, (len(tidx), 1)).astype(np.int64))
x0 = np.ascontiguousarray(truth[0] + rng.standard_normal((B, k, N)))
rho = np.linspace(1.0, 1.5, B)
padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=1.0,
                          obs_error_sd=sd, n_obs_times=len(tidx), n_obs=n_obs, stationary=1,
                          n_cycles=n_cycles, t_pad=1, rank=0, world=1)
H = _native.Handle(0)
mean = np.empty((B, n_cycles, N))
H.etkf_cycle_batched(cfg, B, rho, x0, vals1, idx, 3, padded, None, mean, None)     # warm-up
t0 = time.perf_counter()
H.etkf_cycle_batched(cfg, B, rho, x0, vals1, idx, 3, padded, None, mean, None)
t1 = time.perf_counter()
rmse = np.sqrt(((mean - truth[tidx[:n_cycles]][None]) ** 2).mean(axis=2))[:, 50:].mean(axis=1)
# oracle on 2 experiments
t2 = time.perf_counter()
for b in (0, B - 1):
    X = x0[b]
    w = np.full(n_obs, 1 / sd ** 2)
    for c in range(n_cycles):
        Xa = o_etkf.analysis_sparse(X, vals1[c], idx[c], w, rho[b])
        X = o_l96.rk4_forecast(Xa, S, dt)
t3 = time.perf_counter()
print(json.dumps({"workload": "C2: %d x (N=40,k=20) ETKF, 200 cycles, S=20" % B, "seconds_e2e": t1 - t0,
                  "experiment_cycles_per_s": B * n_cycles / (t1 - t0),
                  "member_steps_per_s": B * n_cycles * k * S / (t1 - t0),
                  "rmse_best_rho": float(rho[np.nanargmin(rmse)]), "rmse_min": float(np.nanmin(rmse)),
                  "cpu_oracle_experiment_cycles_per_s": 2 * n_cycles / (t3 - t2)}))
