"""Debug: section timers of the Newton-Schulz K4 kernel (needs build/libdab_timing.so from tools/build_timing.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dataassimbench_b200 import _native
_native.LIB_PATH = os.path.join(ROOT, 'build', 'libdab_timing.so')
H = _native.Handle(0)
for k, n_y in [(64, 19661), (128, 262144)]:
    rng = np.random.default_rng(0)
    Yp = rng.standard_normal((k, n_y)); Yp -= Yp.mean(axis=0)
    stats = torch.from_numpy(np.concatenate([(Yp @ Yp.T).ravel(), Yp @ rng.standard_normal(n_y)])).cuda()
    T = torch.empty((k, k), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for i in range(3):
        H.etkf_transform_matrix(stats, k, 1.0, False, T, None, None, st)
        torch.cuda.synchronize()
