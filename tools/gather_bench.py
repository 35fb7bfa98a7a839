"""K2 (standalone observation gather, dab_obs_gather_f64): achieved GB/s against its algorithmic
bytes (16 k n_y + 9 n_y: read + write of the gathered values, index and mask) at the bench shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
H = _native.Handle(0)
out = {}
for name, N, k, frac in [("c3", 65536, 64, 0.3), ("c4", 4194304, 64, 0.3), ("c5", 1048576, 128, 1.0)]:
    n_y = int(round(frac * N))
    rng = np.random.default_rng(0)
    idx_np = np.arange(N, dtype=np.int64) if n_y == N else rng.choice(N, n_y, replace=False, shuffle=False).astype(np.int64)
    X = torch.randn(k, N, dtype=torch.float64, device="cuda")
    mask = torch.ones(n_y, dtype=torch.uint8, device="cuda")
    Yb = torch.empty(k, n_y, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for order in ("as drawn", "sorted"):
        idx = torch.from_numpy(np.sort(idx_np) if order == "sorted" else idx_np).cuda()
        for _ in range(3):
            H.obs_gather(X, idx, mask, Yb, N, k, n_y, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            H.obs_gather(X, idx, mask, Yb, N, k, n_y, st)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        assert torch.equal(Yb, X[:, idx])
        alg = 16.0 * k * n_y + 9.0 * n_y
        out["%s %s" % (name, order)] = {"us": round(us, 1), "algorithmic_GBs": round(alg / us / 1e3, 1), "n_y": n_y, "k": k}
        print(name, order, out["%s %s" % (name, order)], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gather_bench.json"), "w"), indent=1)
