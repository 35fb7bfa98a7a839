"""Surface-QG ensemble forecast: device rate (CUDA events on the launching stream) and the NumPy oracle on one host
thread beside it.  N / K (members) / S (steps) / REPS from the environment; prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataassimbench_b200.data import SQGTurb
from dataassimbench_b200 import _runtime

N, k, S, reps = (int(os.environ.get(n, d)) for n, d in (("N", 96), ("K", 64), ("S", 10), ("REPS", 5)))
rng = np.random.default_rng(0)
pv0 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dataassimbench_b200",
                           "_suppl_data", "sqgturb_57600steps.npy")).astype(np.float64)
if N != 96:
    pv0 = 1000.0 * rng.standard_normal((2, N, N))
pv = pv0[None] + 10.0 * rng.standard_normal((k, 2, N, N))
g = SQGTurb(pv=pv[0], precision="double")
H = g._setup_device()
x = torch.from_numpy(np.fft.rfft2(pv)).cuda()
out = torch.empty_like(x)
st = torch.cuda.current_stream()
for _ in range(2):
    H.sqg_forecast(x, out, None, k, S, st.cuda_stream)
torch.cuda.synchronize()
l0 = H.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    H.sqg_forecast(x, out, None, k, S, st.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
M, NK, Nh = 3 * N // 2, N // 2 + 1, N // 2
flop_stage = 2.0 * ((2 * M) * (2 * N) * (8 * NK) + 8 * 2 * M * M * NK + 2 * M * M * N + 2 * 2 * (2 * N) * M * Nh)
flops = 4 * flop_stage * k * S
peak = H.measure_fp64_peak()
# oracle on one host thread: a bounded sample (2 members x S steps)
from oracle import sqgturb as o
T = o.tables(N, "double")
t0 = time.perf_counter()
n_cpu = min(k, 2)
for m in range(n_cpu):
    ref, _, _ = o.integrate(T, np.fft.rfft2(pv[m]), S)
cpu_s = (time.perf_counter() - t0) / n_cpu
err = float(np.abs(out[n_cpu - 1].cpu().numpy() - ref).max() / np.abs(ref).max())
print(json.dumps({"workload": "SQGTurb ensemble forecast", "N": N, "members": k, "steps": S, "ms_per_forecast": ms,
                  "member_steps_per_s": k * S / ms * 1e3, "gemm_tflops": flops / ms / 1e9,
                  "fp64_peak_tflops": peak, "frac_of_dfma_peak": flops / ms / 1e9 / peak["dfma_tflops"],
                  "launches_per_forecast": (H.kernel_launches() - l0) // reps,
                  "oracle_member_steps_per_s_1thread": S / cpu_s, "max_rel_err_vs_oracle": err}))
