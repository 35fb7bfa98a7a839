"""dab_observe_f64 (Observer.observe sampling of a device-resident trajectory): achieved GB/s against
its algorithmic bytes -- per observation 8 (value read) + 8 (error) + 8 (write), plus the index:
8 per observation for moving observers, 8 per location for stationary ones -- at the C4/C5 shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
H = _native.Handle(0)
out = {}
n_t = 10
for name, N, frac in [("c3", 65536, 0.3), ("c4", 4194304, 0.3), ("c5", 1048576, 1.0)]:
    n_obs = int(round(frac * N))
    rng = np.random.default_rng(0)
    state = torch.randn(n_t + 1, N, dtype=torch.float64, device="cuda")
    pos = torch.arange(1, n_t + 1, dtype=torch.int64, device="cuda")
    err = torch.randn(n_t, n_obs, dtype=torch.float64, device="cuda")
    val = torch.empty(n_t, n_obs, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for kind in ("stationary", "moving"):
        if kind == "stationary":
            loc_np = rng.choice(N, n_obs, replace=False, shuffle=False).astype(np.int64) if n_obs < N else np.arange(N, dtype=np.int64)
            cnt = None
        else:
            loc_np = np.stack([rng.choice(N, n_obs, replace=False, shuffle=False) for _ in range(n_t)]).astype(np.int64)
            cnt = torch.full((n_t,), n_obs, dtype=torch.int64, device="cuda")
        loc = torch.from_numpy(loc_np).cuda()
        times = []
        for rep in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            H.observe(state, N, n_t + 1, pos, n_t, loc, kind == "stationary", cnt, n_obs, err, val, st)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(times[2:]))
        ref = (state[1:][:, loc] if kind == "stationary" else torch.gather(state[1:], 1, loc)) + err
        assert torch.equal(val, ref)
        alg = 24.0 * n_t * n_obs + 8.0 * (n_obs if kind == "stationary" else n_t * n_obs)
        out["%s %s" % (name, kind)] = {"us": round(us, 1), "algorithmic_GBs": round(alg / us / 1e3, 1),
                                       "n_obs": n_obs, "n_times": n_t}
        print(name, kind, out["%s %s" % (name, kind)], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "observe_bench.json"), "w"), indent=1)
