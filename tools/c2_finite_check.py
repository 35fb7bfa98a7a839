"""Which experiments of the C2 inflation sweep go non-finite, when, and with which K4 method (debug tool)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from dataassimbench_b200 import _native
H = _native.Handle(0)
rk4, _ = bench.device_generators(H)
N, k, S, n_obs, sd, dt, B, n_cycles = 40, 20, 20, 20, 1.0, 0.01, 4096, 200
rng = np.random.default_rng(1)
truth0 = rk4(8 + rng.standard_normal(N), 1000, dt)
_, truth = rk4(truth0, n_cycles * S, dt, return_traj=True)
tidx = np.arange(0, n_cycles * S + 1, S)
locs = rng.choice(N, n_obs, replace=False, shuffle=False)
vals1 = np.ascontiguousarray(truth[tidx][:, locs] + sd * rng.standard_normal((len(tidx), n_obs)))
idx = np.ascontiguousarray(np.tile(locs, (len(tidx), 1)).astype(np.int64))
x0 = np.ascontiguousarray(truth[0] + rng.standard_normal((B, k, N)))
rho = np.linspace(1.0, 1.5, B)
padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=1.0, obs_error_sd=sd,
                          n_obs_times=len(tidx), n_obs=n_obs, stationary=1, n_cycles=n_cycles, t_pad=1, rank=0, world=1)
for method in (0, 1):
    H.set_eig_method(method)
    mean = np.empty((B, n_cycles, N))
    H.etkf_cycle_batched(cfg, B, rho, x0, vals1, idx, 3, padded, None, mean, None)
    fin = np.isfinite(mean).all(axis=2)            # [B][n_cycles]
    bad = np.where(~fin.all(axis=1))[0]
    first = [int(np.argmin(fin[b])) for b in bad]
    print("method", method, "non-finite experiments", len(bad), "rho range", (rho[bad].min(), rho[bad].max()) if len(bad) else None,
          "first bad cycles", sorted(first)[:10], flush=True)
    good = fin.all(axis=1)
    rm = np.sqrt(np.mean((mean[:, -1, :] - truth[tidx[n_cycles - 1]]) ** 2, axis=1))
    print("  rmse of last mean (finite ones): median %.2f max %.2f" % (np.median(rm[good]), rm[good].max()))
    if len(bad):
        b = bad[0]; c = first[0] if False else int(np.argmin(fin[b]))
        print("  example", b, "rho", rho[b], "mean abs max before:", np.abs(mean[b, max(0, c - 3):c]).max(axis=1))
np.save("gpurun_out/c2_dbg_inputs.npy", np.array([0]))
