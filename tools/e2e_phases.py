"""Host wall-clock phases of the e2e call at C3: setup / H2D + 10 cycles + D2H / whole call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from dataassimbench_b200 import _native
from dataassimbench_b200.sharded import ShardedCycler
w = bench.make_workload("c3")
H = _native.Handle(0)
cyc = ShardedCycler(H, w, bench.BLOCK)
cfg = cyc.config(bench.BLOCK)
k, N = w["k"], w["N"]
x0 = torch.from_numpy(w["X0"]).pin_memory()
out = torch.empty((bench.BLOCK, k, N), dtype=torch.float64).pin_memory()
for _ in range(3):
    H.etkf_cycle_host(cfg, x0, w["vals"], w["sysidx"], w["padded"], out)
def t(f, n=10):
    H.sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    H.sync(); return (time.perf_counter() - t0) / n * 1e3
print("whole call      %.3f ms" % t(lambda: H.etkf_cycle_host(cfg, x0, w["vals"], w["sysidx"], w["padded"], out)))
print("setup only      %.3f ms" % t(lambda: H.cycler_setup(cfg, w["vals"], w["sysidx"], w["padded"])))
def rest():
    H.cycler_set_state(x0, True); H.cycler_run(0, bench.BLOCK, out); H.sync()
print("H2D+run+D2H     %.3f ms" % t(rest))
def run_nod2h():
    H.cycler_set_state(x0, True); H.cycler_run(0, bench.BLOCK); H.sync()
print("H2D+run (no D2H) %.3f ms" % t(run_nod2h))
