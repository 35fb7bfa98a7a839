"""ETKF cycling with the surface-QG model through the public API (`ETKF.cycle()` with `model.SQGTurbModel`): nature run,
seeded Observer on the flattened gridded state, K members, C cycles of W model steps.  Host wall clock around the call
(the ensemble stays on the device between analysis and forecast; observation tables go up once, analyses come back
once).  K / C / W / DENS from the environment; prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dataassimbench_b200 as dab
from dataassimbench_b200 import _xr

k, n_cycles, W = int(os.environ.get("K", 32)), int(os.environ.get("C", 20)), int(os.environ.get("W", 4))
dens, sd = float(os.environ.get("DENS", 0.02)), 100.0
gen = dab.data.SQGTurb(precision="double")
N = gen.N
nat = gen.generate(n_steps=n_cycles * W + 2 * W)
flat = np.asarray(nat["pv"].values, dtype=np.float64).reshape(len(nat["time"].values), -1)
times = np.asarray(nat["time"].values, dtype=np.float64)
nature = _xr.new_dataset({"x": (("time", "index"), flat)}, coords={"time": times, "index": np.arange(flat.shape[1])},
                         attrs={"system_dim": flat.shape[1], "delta_t": 900.0})
obs = dab.observer.Observer(nature, times=times[0::W][:n_cycles], random_location_density=dens, error_sd=sd,
                            random_seed=7).observe()
rng = np.random.default_rng(1)
x0 = flat[0][None] + 300.0 * rng.standard_normal((k, flat.shape[1]))
model = dab.model.SQGTurbModel(model_obj=dab.data.SQGTurb(precision="double"))
dc = dab.dacycler.ETKF(system_dim=flat.shape[1], delta_t=900, ensemble_dim=k, model_obj=model,
                       multiplicative_inflation=1.05)
init = _xr.new_dataset({"x": (("ensemble", "index"), x0)})
best = None
for rep in range(3):
    t0 = time.perf_counter()
    out = dc.cycle(input_state=init, start_time=0.0, obs_vector=obs, obs_error_sd=sd, analysis_window=W * 900.0,
                   n_cycles=n_cycles)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
xa = np.asarray(out["x"].values)
truth = flat[np.arange(n_cycles) * W]                         # a cycle returns its analysis
rmse = np.sqrt(np.mean((xa.mean(axis=1) - truth) ** 2, axis=1))
free = np.sqrt(np.mean((x0.mean(axis=0)[None] - truth) ** 2, axis=1))
print(json.dumps({"workload": "ETKF.cycle() with SQGTurbModel", "N": N, "system_dim": int(flat.shape[1]), "members": k,
                  "n_obs_per_time": int(np.asarray(obs["system_index"].values).shape[-1]), "cycles": n_cycles,
                  "model_steps_per_cycle": W, "seconds_best_of_3": best, "cycles_per_s": n_cycles / best,
                  "member_steps_per_s": n_cycles * k * W / best, "all_finite": bool(np.isfinite(xa).all()),
                  "rmse_first_last": [float(rmse[0]), float(rmse[-1])], "rmse_of_initial_mean_last": float(free[-1])}))
