"""Debug: per-CTA schedule of the forecast kernel (needs build/libdab_timing.so from tools/build_timing.sh).
Prints, for the C3 shape, kernel span, per-phase CTA times and per-SM busy/idle structure."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dataassimbench_b200 import _native
_native.LIB_PATH = os.path.join(ROOT, 'build', 'libdab_timing.so')
H = _native.Handle(0)
lib = _native.load()
N, k, S = int(os.environ.get("N", 65536)), int(os.environ.get("K", 64)), int(os.environ.get("S", 20))
x = (8 + torch.randn(k, N, dtype=torch.float64, device="cuda"))
out = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for shape in os.environ.get("SHAPES", "256,16128,17128,12192,384").split(","):
    os.environ["DAB_L96_T"] = shape
    buf = torch.zeros(8 * 20000, dtype=torch.int64, device="cuda")
    lib.dab_debug_l96_trace.argtypes = [ctypes.c_void_p]
    H.l96_forecast(x, out, None, N, k, S, 0.01, 8.0, st)          # warm
    assert lib.dab_debug_l96_trace(buf.data_ptr()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); H.l96_forecast(x, out, None, N, k, S, 0.01, 8.0, st); e1.record()
    torch.cuda.synchronize()
    assert lib.dab_debug_l96_trace(0) == 0
    b = buf.cpu().numpy().reshape(-1, 8)
    b = b[b[:, 0] > 0]
    t0 = b[:, 0].min()
    start, loaded, comp, end, sm = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3, (b[:, 2] - t0) / 1e3, (b[:, 3] - t0) / 1e3, b[:, 4]
    print("shape %s: %d CTAs, event %.1f us, trace span %.1f us" % (shape, len(b), e0.elapsed_time(e1) * 1e3, end.max()))
    print("  CTA phases (us, median [p5,p95]): load %.2f [%.2f,%.2f]  compute %.2f [%.2f,%.2f]  store %.2f" % (
        np.median(loaded - start), *np.percentile(loaded - start, [5, 95]), np.median(comp - loaded),
        *np.percentile(comp - loaded, [5, 95]), np.median(end - comp)))
    first_start = np.array([start[sm == s].min() for s in np.unique(sm)])
    last_end = np.array([end[sm == s].max() for s in np.unique(sm)])
    n_per_sm = np.array([(sm == s).sum() for s in np.unique(sm)])
    print("  SMs used %d; CTAs/SM min %d max %d; first start: max %.2f us; last end per SM: min %.1f median %.1f max %.1f" % (
        len(first_start), n_per_sm.min(), n_per_sm.max(), first_start.max(), last_end.min(), np.median(last_end), last_end.max()))
    # compute-time by start order deciles (does contention change over the kernel's life?)
    order = np.argsort(start)
    d = np.array_split(order, 8)
    print("  compute time by start-order octile:", " ".join("%.1f" % np.median((comp - loaded)[i]) for i in d))
