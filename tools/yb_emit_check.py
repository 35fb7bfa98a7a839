"""Compact-Yb route (forecast kernels emit the observed columns, dense contraction reads them) against the gather
route: same experiment through two handles (DAB_YB_EMIT=1 / 0), final ensembles compared, cycles timed with CUDA
events on the cycler's stream and the per-kernel profile printed.  N / K / S / DENS / CYCLES from the environment."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native

N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("S", 20))
dens, n_cycles = float(os.environ.get("DENS", 0.3)), int(os.environ.get("CYCLES", 10))
reps = int(os.environ.get("REPS", 20))
n_obs = int(round(dens * N))
rng = np.random.default_rng(0)
vals = 8.0 + rng.standard_normal((n_cycles, n_obs))
sysidx = np.tile(np.sort(rng.choice(N, n_obs, replace=False)), (n_cycles, 1)).astype(np.int64)
padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=0.01, forcing=8.0, rho=1.0, obs_error_sd=1.0,
                          n_obs_times=n_cycles, n_obs=n_obs, stationary=1, n_cycles=n_cycles, t_pad=1, rank=0, world=1)
torch.manual_seed(0)
x0 = 8.0 + torch.randn(k, N, dtype=torch.float64, device="cuda")
res = {}
for emit in (1, 0, 1, 0):
    os.environ["DAB_YB_EMIT"] = str(emit)
    H = _native.Handle(0)
    H.cycler_setup(cfg, vals, sysidx, padded)
    st = torch.cuda.ExternalStream(H.cycler_stream())
    out = torch.empty_like(x0)
    for _ in range(3):
        H.cycler_set_state(x0, False); H.cycler_run(0, n_cycles)
    H.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(reps):
            H.cycler_set_state(x0, False); H.cycler_run(0, n_cycles)
        e1.record()
    H.sync()
    ms = e0.elapsed_time(e1)
    H.cycler_get_state(out, False); H.sync()
    H.set_profiling(True)
    H.cycler_set_state(x0, False); H.cycler_run(0, n_cycles); H.sync()
    prof = H.read_profile()
    H.set_profiling(False)
    res.setdefault(emit, []).append({"cycles_per_s": reps * n_cycles / ms * 1e3, "profile": prof, "shape": H.cycler_forecast_shape()})
    res[("x", emit)] = out.clone()
    H.close()
d = (res[("x", 1)] - res[("x", 0)]).abs().max().item() / res[("x", 0)].abs().max().item()
rec = {"N": N, "k": k, "S": S, "n_obs": n_obs, "cycles": n_cycles, "max_rel_diff_emit_vs_gather": d,
       "emit": res[1], "gather": res[0]}
print(json.dumps(rec, default=str))
