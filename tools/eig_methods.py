"""Time K4 (transform matrix from the statistics) with both solvers on C3/C5-like statistics:
cluster Newton-Schulz (csrc/etkf_ns.cu) vs the one-CTA Jacobi eigensolver (csrc/etkf_eig.cu).
CUDA events on the launching stream, 50 back-to-back launches after warm-up."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200 import _native
from oracle import etkf as o_etkf

H = _native.Handle(0)
s = torch.cuda.Stream()
out = {}
with torch.cuda.stream(s):
    st = s.cuda_stream
    for k, n_y, spread in [(64, 19661, 1.0), (64, 19661, 0.2), (48, 5000, 1.0), (128, 1048576 // 4, 1.0), (96, 30000, 1.0)]:
        rng = np.random.default_rng(0)
        Yp = rng.standard_normal((k, n_y)) * spread; Yp -= Yp.mean(axis=0)
        C, g = Yp @ Yp.T, Yp @ rng.standard_normal(n_y)
        T_ref = o_etkf.transform_from_stats(C, g, k, 1.0)
        stats = torch.from_numpy(np.concatenate([C.ravel(), g])).cuda()
        T = torch.empty((k, k), dtype=torch.float64, device="cuda")
        sw = torch.zeros(1, dtype=torch.int32, device="cuda")
        rec = {}
        for name, m in (("newton_schulz", H.EIG_AUTO), ("jacobi", H.EIG_JACOBI)):
            H.set_eig_method(m)
            for _ in range(5):
                H.etkf_transform_matrix(stats, k, 1.0, False, T, None, sw, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(50):
                H.etkf_transform_matrix(stats, k, 1.0, False, T, None, sw, st)
            e1.record(s); s.synchronize()
            err = float(np.max(np.abs(T.cpu().numpy() - T_ref)) / np.max(np.abs(T_ref)))
            rec[name] = {"us": e0.elapsed_time(e1) * 1e3 / 50, "steps_or_sweeps": int(sw.item()), "rel_err_vs_oracle": err}
        H.set_eig_method(H.EIG_AUTO)
        out["k=%d n_y=%d spread=%g" % (k, n_y, spread)] = rec
        print(k, n_y, spread, rec, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "eig_methods.json"), "w"), indent=1)
