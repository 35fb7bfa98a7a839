"""Three launches of the forecast kernel at the C3 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataassimbench_b200 import _native
N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("S", 20))
H = _native.Handle(0)
x = 8 + torch.randn(k, N, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
