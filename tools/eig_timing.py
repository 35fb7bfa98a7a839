"""Debug: section timers of the block-Jacobi eigensolver (needs build/libdab_timing.so, built with
-DDAB_EIG_TIMING by tools/build_timing.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dataassimbench_b200 import _native
_native.LIB_PATH = os.path.join(ROOT, 'build', 'libdab_timing.so')
k, n_y = int(os.environ.get("K", 64)), int(os.environ.get("NY", 19661))
rng = np.random.default_rng(0)
Yp = rng.standard_normal((k, n_y)); Yp -= Yp.mean(axis=0)
stats = torch.from_numpy(np.concatenate([(Yp @ Yp.T).ravel(), Yp @ rng.standard_normal(n_y)])).cuda()
H = _native.Handle(0)
T = torch.empty((k, k), dtype=torch.float64, device="cuda")
lam = torch.zeros(k, dtype=torch.float64, device="cuda")
sw = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for i in range(2):
    H.etkf_transform_matrix(stats, k, 1.0, False, T, lam, sw, st)
torch.cuda.synchronize()
l = lam.cpu().numpy()[:6]
print("sweeps", int(sw.item()), dict(zip(["phaseA", "load", "sub4", "store_sync", "flags", "build"], l.tolist())), "total", l.sum())
