"""Diagnostic: multi-cycle runs with programmatic dependent launch off / on, per-cycle differences."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_gpu_kernels as tk
from dataassimbench_b200._runtime import handle

H = handle()
for (N, k, n_obs) in [(36, 10, 18), (40, 20, 20), (6000, 64, 1800), (3000, 128, 3000)]:
    n_cycles, S, sd, rho = 10, 20, 1.0, 1.03
    case = tk.cycler_case(N, k, n_obs, n_cycles, S, sd, rho, True, seed=N + 3 * k)
    runs = []
    for pdl in (0, 0, 1, 1, 0):
        H.set_pdl(pdl)
        out, final, _ = tk.run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
        runs.append((pdl, out, final))
    base = runs[0]
    for i, (pdl, out, final) in enumerate(runs[1:], 1):
        d = np.abs(out - base[1]).reshape(n_cycles, -1).max(axis=1)
        print(N, k, 'run', i, 'pdl', pdl, 'max diff per cycle', ' '.join('%.1e' % v for v in d),
              'final', np.abs(final - base[2]).max(), flush=True)
H.set_pdl(1)
