"""Experiment: forecast kernel time vs CTA width (env DAB_L96_T) at the C3 shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataassimbench_b200 import _native

N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), 20
H = _native.Handle(0)
x = (8 + torch.randn(k, N, dtype=torch.float64, device="cuda"))
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
res = {}
for T in [0] + list(range(96, 513, 32)):
    if T:
        os.environ["DAB_L96_T"] = str(T)
    else:
        os.environ.pop("DAB_L96_T", None)
    for _ in range(3):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    res[T] = round(us, 1)
print(json.dumps(res))
