"""Experiment: forecast kernel time vs CTA shape at one problem shape (default C3), through the
standalone entry point with the shape forced by DAB_L96_T (see kernels.h: l96_shape_code).
The cycler does the same measurement itself at setup (engine.cu; DAB_TUNE_VERBOSE=1 prints it)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataassimbench_b200 import _native

N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("S", 20))
H = _native.Handle(0)
x = (8 + torch.randn(k, N, dtype=torch.float64, device="cuda"))
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
res = {}
ref = None
shapes = [0, 128, 192, 256, 320, 384, 512, 12128, 12192, 16128, 16192, 16256, 17128, 17160, 17192]
for code in shapes:
    os.environ.pop("DAB_L96_T", None)
    if code:
        os.environ["DAB_L96_T"] = str(code)
    for _ in range(3):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    res[str(code)] = (round(e0.elapsed_time(e1) / 20 * 1e3, 1), bool(torch.equal(ref, y)))
print(json.dumps(res))
