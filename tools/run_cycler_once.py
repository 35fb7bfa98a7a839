"""One cycler set-up and two cycles at a given shape (for ncu: the forecast kernel the cycler really launches).
N / K / NOBS / S from the environment; DAB_L96_V6 / DAB_L96_V3 / DAB_L96_T force a forecast kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("S", 20))
n_obs, n_cycles = int(os.environ.get("NOBS", 4096)), 2
H = _native.Handle(0)
rng = np.random.default_rng(0)
vals = 8.0 + rng.standard_normal((n_cycles, n_obs))
sysidx = np.tile(np.sort(rng.choice(N, n_obs, replace=False)), (n_cycles, 1)).astype(np.int64)
padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=0.01, forcing=8.0, rho=1.0, obs_error_sd=1.0,
                          n_obs_times=n_cycles, n_obs=n_obs, stationary=1, n_cycles=n_cycles, t_pad=1, rank=0, world=1)
H.cycler_setup(cfg, vals, sysidx, padded)
H.cycler_set_state(8.0 + torch.randn(k, N, dtype=torch.float64, device="cuda"), False)
H.cycler_run(0, n_cycles)
H.sync()
print("ok", H.cycler_forecast_shape())
