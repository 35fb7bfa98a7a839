"""Three launches of the eigensolver/transform-matrix kernel on C3-like statistics (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
k, n_y = int(os.environ.get("K", 64)), int(os.environ.get("NY", 19661))
rng = np.random.default_rng(0)
Yp = rng.standard_normal((k, n_y)); Yp -= Yp.mean(axis=0)
stats = torch.from_numpy(np.concatenate([(Yp @ Yp.T).ravel(), Yp @ rng.standard_normal(n_y)])).cuda()
H = _native.Handle(0)
T = torch.empty((k, k), dtype=torch.float64, device="cuda")
sw = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record()
    H.etkf_transform_matrix(stats, k, 1.0, False, T, None, sw, st)
    e1.record(); torch.cuda.synchronize()
print("ok sweeps", int(sw.item()), "us", e0.elapsed_time(e1) * 1e3)
