"""Experiment: forecast kernel time vs kernel variant / CTA width at the C3 shape.
v1 = per-stage shared-memory exchange (env DAB_L96_T), v2 = persistent + shuffle (env DAB_L96_V2)."""
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
for var, T in [("v1", 0), ("v1", 256), ("v1", 512)] + [("v2", t) for t in range(128, 513, 64)]:
    os.environ.pop("DAB_L96_T", None); os.environ.pop("DAB_L96_V2", None)
    if T:
        os.environ["DAB_L96_T" if var == "v1" else "DAB_L96_V2"] = str(T)
    for _ in range(3):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        H.l96_forecast(x, y, None, N, k, S, 0.01, 8.0, st)
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    same = bool(torch.equal(ref, y))
    res["%s_%d" % (var, T)] = (round(e0.elapsed_time(e1) / 20 * 1e3, 1), same)
print(json.dumps(res))
