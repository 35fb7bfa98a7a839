"""A few single analyses (K3 contraction, reduce, K4, K5 transform) at the C3 shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
N, k, n_y = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("NY", 19661))
H = _native.Handle(0)
rng = np.random.default_rng(0)
X = 8 + torch.randn(k, N, dtype=torch.float64, device="cuda")
Xa = torch.empty_like(X)
idx = torch.from_numpy(np.sort(rng.choice(N, n_y, replace=False)).astype(np.int64)).cuda()
y = 8 + torch.randn(n_y, dtype=torch.float64, device="cuda")
mask = torch.ones(n_y, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    H.etkf_analysis(X, Xa, y, idx, mask, 1.0, 1.0, N, k, n_y, st)
torch.cuda.synchronize()
print("ok", float(Xa.abs().max()))
