#!/usr/bin/env python
"""bench.py -- ETKF cycles/s on Lorenz-96 N=65,536, k=64 (BASELINE.json configs[2], "C3").

One "step" = one ETKF cycle = analysis (gather + contraction + eigensolver + transform) followed
by the S=20-step RK4 ensemble forecast.  Prints ONE JSON line (see the task contract):

  value          cycles/s with everything resident in HBM (CUDA events on the launching stream)
  e2e            cycles/s through the host-buffer C-ABI call `dab_etkf_cycle_host_f64`
                 (H2D of ensemble + observations and D2H of every analysis inside the timed region)
  roofline       the forecast kernel (FP64-pipe bound) against the FP64 peak measured in this run
  kernels        time share and roofline of every kernel of the cycle
  cpu_baseline   the NumPy oracle (a restatement of the reference, NOT JAX) on this box's host cores

`--impl reference` times only the CPU oracle ("port": JAX/xarray are not installable here).

The filter diverges after ~20 cycles at these shapes (k << N, no localisation), so
the experiment is a 10-cycle run from a fresh ensemble, repeated; the reset (one 33.5 MB D2D copy
per 10 cycles) stays INSIDE the timed region, and the final state is asserted finite.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, k, obs fraction, S, dt, sd, rho
    "c3": dict(N=65536, k=64, obs_frac=0.3, S=20, dt=0.01, sd=1.0, rho=1.0,
               label="Lorenz96 N=65536 k=64 ETKF, 30% random stationary obs, S=20 RK4 steps/cycle"),
    "c4": dict(N=4194304, k=64, obs_frac=0.3, S=20, dt=0.01, sd=1.0, rho=1.0,
               label="Lorenz96 N=4194304 k=64 ETKF, 30% random stationary obs, S=20"),
    "c5": dict(N=1048576, k=128, obs_frac=1.0, S=20, dt=0.01, sd=1.0, rho=1.0,
               label="Lorenz96 N=1048576 k=128 ETKF, 100% obs, S=20"),
    "c1": dict(N=36, k=10, obs_frac=0.5, S=20, dt=0.01, sd=1.0, rho=1.0,
               label="Lorenz96 N=36 k=10 ETKF, 50% stationary obs, S=20"),
}
BLOCK = 10          # cycles per experiment (fresh ensemble every BLOCK cycles)
FLOPS_PER_POINT_STEP = 29.0


def oracle_generators():
    """(forecast, padded_table) built on the CPU oracle: the reference arm, the cpu_baseline leg and
    the offline tools.  The product arm never calls this (see device_generators)."""
    from oracle import l96 as o_l96, windows as o_win

    def forecast(x, n, dt, return_traj=False):
        return o_l96.rk4_forecast(x, n, dt, return_traj=return_traj)

    def padded_table(obs_times, window, n_cycles):
        times = o_win.get_all_times(0.0, window, n_cycles)
        return np.ascontiguousarray(o_win.pad_time_indices(o_win.get_obs_indices(times, obs_times, window)))

    return forecast, padded_table


def device_generators(H):
    """The same two functions made of the package itself: the nature run comes from the forecast
    kernel (dab_l96_forecast_f64) and the window table from the host mirror of the reference's
    bookkeeping, so the product arm's synthetic inputs involve nothing under oracle/."""
    import torch
    from dataassimbench_b200.dacycler import _windows

    def forecast(x, n, dt, return_traj=False):
        x = np.asarray(x, dtype=np.float64)
        a = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(x))).cuda()
        b = torch.empty_like(a)
        tr = torch.empty((int(n),) + tuple(a.shape), dtype=torch.float64, device="cuda") if return_traj else None
        H.l96_forecast(a, b, tr, a.shape[1], a.shape[0], int(n), dt, 8.0)
        torch.cuda.synchronize()
        last = b.cpu().numpy().reshape(x.shape)
        if not return_traj:
            return last
        full = np.concatenate([np.atleast_2d(x)[None], tr.cpu().numpy()], axis=0)
        return last, full.reshape((int(n) + 1,) + x.shape)

    def padded_table(obs_times, window, n_cycles):
        centres = _windows.window_centres(0.0, window, n_cycles)
        return np.ascontiguousarray(_windows.padded_time_table(obs_times, centres, window))

    return forecast, padded_table


def make_workload(name, seed=42, generators=None, n_cycles=None):
    """Synthetic experiment: truth spun up with RK4, ensemble = truth + N(0,1),
    one observation time per window placed at the analysis time, stationary random locations
    drawn like Observer does (default_rng.choice(N, n_obs, replace=False, shuffle=False))."""
    rk4, padded_table = generators if generators is not None else oracle_generators()
    w = dict(WORKLOADS[name])
    N, k, S, dt = w["N"], w["k"], w["S"], w["dt"]
    n_cycles = BLOCK if n_cycles is None else int(n_cycles)
    rng = np.random.default_rng(seed)
    truth = 8.0 + rng.standard_normal(N)
    # spin-up onto the attractor (local decorrelation takes a few time units); bounded so that the
    # CPU-side generation of the multi-million-point workloads stays within a minute
    truth = rk4(truth, int(max(100, min(500, 6.0e7 / N))), dt)
    n_obs = int(round(w["obs_frac"] * N))
    locs = np.sort(rng.choice(N, n_obs, replace=False, shuffle=False)) if n_obs == N else \
        rng.choice(N, n_obs, replace=False, shuffle=False)
    X0 = truth + rng.standard_normal((k, N))
    vals = np.empty((n_cycles, n_obs))
    x = truth
    for c in range(n_cycles):
        vals[c] = x[locs] + w["sd"] * rng.standard_normal(n_obs)
        if c + 1 < n_cycles:
            x = rk4(x, S, dt)
    obs_times = np.arange(n_cycles) * (S * dt)
    padded = padded_table(obs_times, S * dt, n_cycles)
    w.update(n_obs=n_obs, X0=np.ascontiguousarray(X0), vals=np.ascontiguousarray(vals),
             sysidx=np.ascontiguousarray(np.tile(locs, (n_cycles, 1)).astype(np.int64)), padded=padded,
             name=name, obs_times=obs_times, n_cycles=n_cycles)
    return w


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------ CPU oracle
_POOL = None


def _forecast_all_cores(X, S, dt):
    """Oracle RK4 over the ensemble, member chunks on all host cores (NumPy ufuncs release the GIL)."""
    global _POOL
    from oracle import l96 as o_l96
    n = min(os.cpu_count() or 1, X.shape[0])
    if n <= 1:
        return o_l96.rk4_forecast(X, S, dt)
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=os.cpu_count() or 1)
    chunks = np.array_split(np.arange(X.shape[0]), n)
    parts = list(_POOL.map(lambda ix: o_l96.rk4_forecast(X[ix[0]:ix[-1] + 1], S, dt), chunks))
    return np.concatenate(parts, axis=0)


def oracle_cycles(w, n_cycles):
    """Run n_cycles cycles of the ref-sparse oracle (analysis_sparse: BLAS threads; RK4: member
    chunks on every host core)."""
    from oracle import etkf as o_etkf
    X = w["X0"].copy()
    wts = np.full(w["n_obs"], 1.0 / w["sd"] ** 2)
    for c in range(n_cycles):
        cc = c % BLOCK
        if cc == 0:
            X = w["X0"].copy()
        Xa = o_etkf.analysis_sparse(X, w["vals"][cc], w["sysidx"][cc], wts, w["rho"])
        X = _forecast_all_cores(Xa, w["S"], w["dt"])
    return X


def cpu_threads():
    """Host threads the CPU arm uses: the forecast pool spans every core; BLAS reports its own."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] + [n])
    except Exception:
        pass
    return int(n)


def time_oracle(w, steps, warmup, budget_s=60.0):
    """(cycles/s, cycles timed).  Bounded: stops early once `budget_s` is spent."""
    for _ in range(warmup):
        oracle_cycles(w, 1)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        oracle_cycles(w, 1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """nvidia-smi's clocks line through NVML (same counters), every 20 ms."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                pw = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                self.samples.append((time.perf_counter(), sm, int(reasons), pw))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self, t0, t1):
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting",
                 0x100: "display_clocks_setting"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if len(inside) < 3:
            inside, window = self.samples, "whole-run"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        active = set()
        for s in inside:
            for bit, n in names.items():
                if s[2] & bit:
                    active.add(n)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_sm,
                "reasons": sorted(active), "samples": len(inside), "window": window,
                "power_w_max": max(s[3] for s in inside)}


# ------------------------------------------------------------------------------ reference arm
def shared_config(w):
    """The keys both arms print under `config` (the driver compares them)."""
    return {"workload": w["label"], "N": w["N"], "k": w["k"], "n_obs": w["n_obs"], "S": w["S"],
            "window": w["S"] * w["dt"], "integrator": "rk4"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload)
    cps, done = time_oracle(w, args.steps, args.warmup, budget_s=args.cpu_budget)
    k, S = w["k"], w["S"]
    line = {
        "impl": "reference", "metric": "etkf_cycles_per_s", "value": cps, "unit": "cycles/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1e3 / cps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(w),
        "member_steps_per_s": cps * k * S,
        "cpu_baseline": {"value": cps, "unit": "cycles/s", "cores": cpu_threads(), "kind": "port",
                         "sample": "%d full cycles of the NumPy/SciPy oracle (gather + diagonal R + eigh + "
                                   "RK4 over member chunks on all cores); JAX/xarray are absent, so this is a restatement of "
                                   "the reference, not the reference" % done,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": cps, "unit": "cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_c2_batched(H, B=4096, n_cycles=200):
    """4096 independent Lorenz-96 N=40, k=20 ETKF experiments (inflation sweep 1.0..1.5), S=20,
    through dab_etkf_cycle_batched_f64 with host buffers; wall clock of the second call."""
    from dataassimbench_b200 import _native
    rk4, _ = device_generators(H)
    N, k, S, n_obs, sd, dt = 40, 20, 20, 20, 1.0, 0.01
    rng = np.random.default_rng(1)
    truth0 = rk4(8 + rng.standard_normal(N), 1000, dt)
    _, truth = rk4(truth0, n_cycles * S, dt, return_traj=True)
    tidx = np.arange(0, n_cycles * S + 1, S)
    locs = rng.choice(N, n_obs, replace=False, shuffle=False)
    vals1 = truth[tidx][:, locs] + sd * rng.standard_normal((len(tidx), n_obs))
    vals1 = np.ascontiguousarray(vals1)          # one observation set for the whole sweep (shared_index = 3)
    idx = np.ascontiguousarray(np.tile(locs, (len(tidx), 1)).astype(np.int64))
    x0 = np.ascontiguousarray(truth[0] + rng.standard_normal((B, k, N)))
    rho = np.linspace(1.0, 1.5, B)
    padded = np.arange(1, n_cycles + 1, dtype=np.int64)[:, None].copy()
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=1.0,
                              obs_error_sd=sd, n_obs_times=len(tidx), n_obs=n_obs, stationary=1,
                              n_cycles=n_cycles, t_pad=1, rank=0, world=1)
    mean = np.empty((B, n_cycles, N))
    H.etkf_cycle_batched(cfg, B, rho, x0, vals1, idx, 3, padded, None, mean, None)
    t0 = time.perf_counter()
    H.etkf_cycle_batched(cfg, B, rho, x0, vals1, idx, 3, padded, None, mean, None)
    t1 = time.perf_counter()
    finite = np.isfinite(mean).all(axis=(1, 2))
    return {"workload": "%d x (Lorenz96 N=40, k=20) ETKF, inflation sweep %.2f..%.2f, %d cycles, S=20, one launch"
                        % (B, rho[0], rho[-1], n_cycles),
            "experiment_cycles_per_s": B * n_cycles / (t1 - t0), "member_steps_per_s": B * n_cycles * k * S / (t1 - t0),
            "seconds": t1 - t0, "finite_experiments": int(finite.sum()),
            "largest_inflation_still_finite": float(rho[finite].max()) if finite.any() else None,
            "smallest_inflation_not_finite": float(rho[~finite].min()) if (~finite).any() else None,
            "timer": "host wall clock, host buffers in, ensemble-mean analyses out"}


def run_sqgturb(H, peaks, members=64, n_steps=10, reps=3):
    """the surface-QG generator behind the forecast plugin point.  64 perturbed copies of the
    reference's default state (N = 96), 10 RK4 steps per call; device-resident spectral states, CUDA events."""
    import torch
    from dataassimbench_b200.data import SQGTurb
    g = SQGTurb(precision="double")
    N = g.N
    rng = np.random.default_rng(3)
    pv = g.x0_gridded[None].astype(np.float64) + 10.0 * rng.standard_normal((members, 2, N, N))
    Hs = g._setup_device()
    x = torch.from_numpy(np.fft.rfft2(pv)).cuda()
    out = torch.empty_like(x)
    st = torch.cuda.current_stream()
    Hs.sqg_forecast(x, out, None, members, n_steps, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        Hs.sqg_forecast(x, out, None, members, n_steps, st.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    M, NK, Nh = 3 * N // 2, N // 2 + 1, N // 2
    flop_stage = 2.0 * ((2 * M) * (2 * N) * (8 * NK) + 8 * 2 * M * M * NK + 2 * M * M * N + 2 * 2 * (2 * N) * M * Nh)
    tflops = 4 * flop_stage * members * n_steps / ms / 1e9
    return {"workload": "SQGTurb N=%d (3N/2 dealiasing grid), %d members, %d RK4 steps per call" % (N, members, n_steps),
            "member_steps_per_s": members * n_steps / ms * 1e3, "ms_per_call": ms,
            "dft_gemm_tflops": tflops, "frac_fp64": tflops / max(peaks["dmma_tflops"], peaks["dfma_tflops"]),
            "all_finite": bool(torch.isfinite(torch.view_as_real(out)).all().item()),
            "timer": "CUDA events on the launching stream, states resident in HBM"}


# ------------------------------------------------------------------------------ product arm
def load_peaks():
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return measured.get("hbm_gbs", 6650.0), ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in measured else "fallback")


def kernel_table(prof, w, world, fp64_peak, hbm_peak):
    """Per-kernel time share and roofline fractions from the CUDA-event profiling pass; algorithmic
    flops / bytes per launch on one rank as in SURVEY.md 8d / DESIGN.md 4."""
    N, k, S = w["N"], w["k"], w["S"]
    n_loc = N // world
    n_y = w["n_obs"] / world
    alg = {
        "forecast": dict(flops=FLOPS_PER_POINT_STEP * n_loc * k * S, bytes=16.0 * n_loc * k),
        "transform": dict(flops=2.0 * n_loc * k * k, bytes=16.0 * n_loc * k),
        "contraction": dict(flops=2.0 * n_y * k * k + 4.0 * n_y * k, bytes=8.0 * k * n_y + 16.0 * n_y),
        "reduction": dict(flops=0.0, bytes=0.0),
        "eigensolver": dict(flops=0.0, bytes=0.0),
    }
    kernels = {}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    for name, (tms, cnt) in prof.items():
        if cnt == 0:
            continue
        us = 1e3 * tms / cnt
        ent = {"us": round(us, 2), "share": round(tms / tot_ms, 4)}
        a = alg[name]
        if a["flops"]:
            ent["tflops"] = round(a["flops"] / (us * 1e-6) / 1e12, 3)
            ent["frac_fp64"] = round(ent["tflops"] / fp64_peak, 4)
        if a["bytes"]:
            ent["gbs"] = round(a["bytes"] / (us * 1e-6) / 1e9, 1)
            ent["frac_hbm"] = round(ent["gbs"] / hbm_peak, 4)
        kernels[name] = ent
    return kernels, alg


class Ctx:
    """Rank plumbing shared by the workload runs of one bench invocation."""

    def __init__(self, torch, dist, rank, world, local_rank):
        self.torch, self.dist, self.rank, self.world, self.local_rank = torch, dist, rank, world, local_rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        if self.world == 1:
            return bool(ok)
        f = self.torch.tensor([1.0 if ok else 0.0], device="cuda")
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MIN)
        return bool(f.item() > 0.5)


def timed_cycles(ctx, H, cyc, K, W):
    """W untimed + K timed cycles (experiments of BLOCK cycles from a fresh ensemble, reset inside the
    timed region); CUDA events on the cycler's stream, max over ranks.  Returns (ms, launches, t0, t1)."""
    torch = ctx.torch
    stream = torch.cuda.ExternalStream(H.cycler_stream())

    def run_cycles(n):
        done = 0
        while done < n:
            m = min(BLOCK, n - done)
            cyc.reset()
            cyc.run(0, m)
            done += m

    run_cycles(W)
    H.sync()
    ctx.barrier()
    launches0 = H.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    run_cycles(K)
    ev1.record(stream)
    H.sync()
    ctx.barrier()
    t_wall1 = time.perf_counter()
    ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    return ms, H.kernel_launches() - launches0, t_wall0, t_wall1, run_cycles


def profile_kernels(H, run_cycles, K):
    """Per-kernel timing in a separate pass (event records between kernels perturb the main timing)."""
    H.set_profiling(True)
    run_cycles(BLOCK)            # warm
    H.read_profile()
    run_cycles(max(BLOCK, min(K, 100)))
    prof = H.read_profile()
    H.set_profiling(False)
    return prof


def single_gpu_reference(ctx, w, K):
    """Rank 0 alone runs the same experiment on ONE GPU through a second handle: its rate (for the strong-
    scaling efficiency of this workload) and the state after BLOCK cycles (for the N-rank parity check)."""
    from dataassimbench_b200 import _native
    from dataassimbench_b200.sharded import ShardedCycler
    H1 = _native.Handle(ctx.local_rank)
    one = Ctx(ctx.torch, ctx.dist, 0, 1, ctx.local_rank)
    cyc1 = ShardedCycler(H1, w, BLOCK, 0, 1)
    ms, _, _, _, _ = timed_cycles(one, H1, cyc1, K, BLOCK)
    cyc1.reset()
    cyc1.run(0, 1)
    first = cyc1.state()
    cyc1.run(1, BLOCK)
    final = cyc1.state()
    H1.sync()
    del cyc1
    H1.close()
    return K / (ms * 1e-3), (first, final)


def parity_vs_single_gpu(ctx, cyc, final1):
    """max |x_N - x_1| / max |x_1| of the ensemble after BLOCK cycles from the same fresh ensemble: the
    N-rank run (slabs gathered on rank 0 over NCCL) against rank 0's single-GPU run."""
    torch, dist = ctx.torch, ctx.dist
    out = []
    cyc.reset()
    for i, (c0, c1) in enumerate(((0, 1), (1, BLOCK))):
        cyc.run(c0, c1)
        slab = cyc.state().contiguous()
        parts = [torch.empty_like(slab) for _ in range(ctx.world)] if ctx.rank == 0 else None
        dist.gather(slab, parts, dst=0)
        if ctx.rank == 0:
            full = torch.cat(parts, dim=1)
            out.append(float(((full - final1[i]).abs().max() / final1[i].abs().max()).item()))
            del full, parts
    return out if ctx.rank == 0 else None


def run_workload(ctx, H, name, K, W, peaks, hbm):
    """One workload on all ranks: rate, per-kernel table, forecast roofline; with more than one rank also the
    single-GPU rate measured in the same run and the parity of the N-rank result against it."""
    from dataassimbench_b200.sharded import ShardedCycler
    w = make_workload(name, generators=device_generators(H))
    cps1 = final1 = None
    if ctx.world > 1:
        if ctx.rank == 0:
            cps1, final1 = single_gpu_reference(ctx, w, max(BLOCK, min(K, 5 * BLOCK)))
        ctx.barrier()
    cyc = ShardedCycler(H, w, BLOCK, ctx.rank, ctx.world)
    ms, launches, t0, t1, run_cycles = timed_cycles(ctx, H, cyc, K, W)
    final_ok = ctx.all_true(cyc.state_is_finite())
    prof = profile_kernels(H, run_cycles, K)
    parity = parity_vs_single_gpu(ctx, cyc, final1) if ctx.world > 1 else None
    fp64_peak = max(peaks["dmma_tflops"], peaks["dfma_tflops"])
    kernels, alg = kernel_table(prof, w, ctx.world, fp64_peak, hbm[0])
    cps = K / (ms * 1e-3)
    rec = {"workload": w["label"], "cycles_per_s": cps, "ms_per_cycle": ms / K, "steps": K, "warmup": W,
           "member_steps_per_s": cps * w["k"] * w["S"], "all_finite": final_ok,
           "forecast_frac_fp64": kernels.get("forecast", {}).get("frac_fp64"),
           "forecast_shape": dict(zip(("code", "launches_per_window"), H.cycler_forecast_shape())),
           "kernels": kernels}
    if ctx.world > 1:
        rec["single_gpu_cycles_per_s"] = cps1
        rec["efficiency_vs_1gpu"] = None if cps1 is None else cps / (ctx.world * cps1)
        rec["parity_vs_1gpu"] = None if parity is None else parity[0]
        rec["parity_vs_1gpu_after_%d_cycles" % BLOCK] = None if parity is None else parity[1]
        rec["parity_note"] = ("max |x_N - x_1| / max |x_1| over the whole ensemble after 1 cycle (analysis + forecast; the north "
                              "star's single-cycle tolerance is 1e-10) and after %d cycles from the same fresh ensemble (the "
                              "statistics are summed in a different order on N ranks; the chaotic dynamics double that "
                              "rounding difference every ~2 cycles); the 1-GPU run is rank 0's, in this process, through a "
                              "second handle" % BLOCK)
    return rec, dict(w=w, cyc=cyc, ms=ms, launches=launches, t0=t0, t1=t1, kernels=kernels, alg=alg)


def run_e2e(ctx, H, w, cyc, K):
    """Cycles/s with HOST buffers in and out (see the module docstring)."""
    torch, dist = ctx.torch, ctx.dist
    N, k = w["N"], w["k"]
    reps = max(1, min(K, 200) // BLOCK)
    if ctx.world == 1:
        x0 = torch.from_numpy(w["X0"]).pin_memory()
        out = torch.empty((BLOCK, k, N), dtype=torch.float64).pin_memory()
        cfg = cyc.config(n_cycles=BLOCK)
        H.etkf_cycle_host(cfg, x0, w["vals"], w["sysidx"], w["padded"], out)      # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            H.etkf_cycle_host(cfg, x0, w["vals"], w["sysidx"], w["padded"], out)
        t1 = time.perf_counter()
        assert bool(torch.isfinite(out).all())
        h2d = (w["X0"].nbytes + w["vals"].nbytes + w["sysidx"].nbytes + w["padded"].nbytes) / BLOCK
        return {"value": reps * BLOCK / (t1 - t0), "unit": "cycles/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(k * N * 8),
                "timer": "host wall clock around dab_etkf_cycle_host_f64 (setup, H2D, %d cycles, "
                         "per-cycle D2H of the analysis, sync); ensemble and result buffers pinned" % BLOCK,
                "steps": reps * BLOCK}
    # grid-sharded: every rank uploads its slab of the ensemble from pinned host memory, cycles, and copies
    # its slab of EVERY analysis back to pinned host memory over its own PCIe link (copy stream, overlapped
    # with the next cycle); observations were uploaded at setup.
    stream = torch.cuda.ExternalStream(H.cycler_stream())
    lo, hi = cyc.plan.slab(ctx.rank)
    x0 = torch.from_numpy(np.ascontiguousarray(w["X0"][:, lo:hi])).pin_memory()
    out_host = torch.empty((BLOCK, k, hi - lo), dtype=torch.float64).pin_memory()

    def one_block():
        H.cycler_set_state(x0, True)
        with torch.cuda.stream(stream):
            for c in range(BLOCK):
                H.cycler_stats(c)
                if not cyc.device_exchange:
                    dist.all_reduce(cyc.stats)
                H.cycler_update_host(c, out_host[c])      # D2H on the handle's copy stream
        H.sync()

    one_block()                                                     # warm-up
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        one_block()
    ctx.barrier()
    t1 = time.perf_counter()
    tt = ctx.max_over_ranks(t1 - t0)
    assert ctx.all_true(bool(torch.isfinite(out_host).all()))
    return {"value": reps * BLOCK / tt, "unit": "cycles/s",
            "h2d_bytes_per_step": int(w["X0"].nbytes / BLOCK), "d2h_bytes_per_step": int(k * N * 8),
            "timer": "host wall clock, max over ranks: per rank H2D of its ensemble slab, %d cycles, per-cycle "
                     "D2H of its slab of the analysis (copy stream), sync; bytes are the whole job's" % BLOCK,
            "steps": reps * BLOCK}


def run_e2e_api(w, K):
    """The same metric through the public Python API, `dacycler.ETKF(...).cycle()`, with pageable NumPy inputs
    in a Dataset and a Dataset of analyses out: what a dabench user's script does (N = 1)."""
    import dataassimbench_b200 as dab
    from dataassimbench_b200 import _xr
    N, k, S, dt = w["N"], w["k"], w["S"], w["dt"]
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=dt, store_as_jax=True))
    dc = dab.dacycler.ETKF(system_dim=N, delta_t=dt, ensemble_dim=k, model_obj=model,
                           multiplicative_inflation=w["rho"])
    init = _xr.new_dataset({"x": (("ensemble", "index"), w["X0"])}, coords={"time": 0.0})
    obs = _xr.new_dataset({"x": (("time", "observations"), w["vals"]),
                           "system_index": (("variable", "time", "observations"), w["sysidx"][None])},
                          coords={"time": w["obs_times"], "variable": np.array(["x"])},
                          attrs={"error_sd": w["sd"], "stationary_observers": True})
    kw = dict(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=BLOCK, obs_error_sd=w["sd"],
              analysis_window=S * dt, analysis_time_in_window=0.0)
    out = dc.cycle(**kw)                                            # warm-up: two calls, so that torch's caching host
    out = dc.cycle(**kw)                                            # allocator holds both result blocks the loop alternates between
    reps = max(1, min(K, 100) // BLOCK)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = dc.cycle(**kw)
    t1 = time.perf_counter()
    assert bool(np.isfinite(out["x"].values[-1]).all())
    return {"value": reps * BLOCK / (t1 - t0), "unit": "cycles/s", "steps": reps * BLOCK,
            "timer": "host wall clock around ETKF.cycle(): window bookkeeping, setup, H2D from pageable NumPy, "
                     "%d cycles, per-cycle D2H, Dataset out" % BLOCK}


def run_product(args):
    import torch
    import torch.distributed as dist
    from dataassimbench_b200 import _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)"
                         % (args.gpus, world))
    ctx = Ctx(torch, dist, rank, world, local_rank)
    H = _native.Handle(local_rank)
    peaks = H.measure_fp64_peak()
    hbm = load_peaks()
    K, W = args.steps, args.warmup

    sampler = ClockSampler(local_rank)
    sampler.start()
    rec, aux = run_workload(ctx, H, args.workload, K, W, peaks, hbm)
    w, cyc = aux["w"], aux["cyc"]
    e2e = run_e2e(ctx, H, w, cyc, K)
    sampler.stop_flag = True
    clocks = sampler.summary(aux["t0"], aux["t1"])

    # ---- the other BASELINE configs (sub-records of the same line; the headline stays `args.workload`)
    extra = {}
    if w["name"] == "c3" and not args.no_extra:
        del cyc
        for name in ("c4", "c5"):
            r2, aux2 = run_workload(ctx, H, name, max(BLOCK, min(K, 3 * BLOCK)), BLOCK, peaks, hbm)
            del aux2
            extra[name] = r2
        if world == 1:
            extra["c2_batched"] = run_c2_batched(H, n_cycles=200)
            c1, _ = run_workload(ctx, H, "c1", 200, 20, peaks, hbm)
            extra["c1"] = {kk: c1[kk] for kk in ("workload", "cycles_per_s", "ms_per_cycle", "steps", "all_finite")}
            extra["e2e_api"] = run_e2e_api(w, K)
            extra["sqgturb"] = run_sqgturb(H, peaks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_peak = max(peaks["dmma_tflops"], peaks["dfma_tflops"])
    fc = rec["kernels"].get("forecast", {})
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("forecast")
    except Exception:
        pass
    roofline = {"kernel": "forecast (K1; kernel chosen by the setup-time tuner: config_detail.forecast_shape)",
                "bound": "fp64", "achieved": fc.get("tflops"), "peak": round(fp64_peak, 2), "unit": "TFLOP/s",
                "frac": fc.get("frac_fp64"), "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                  "ncu --set full capture of this kernel at this workload (cold cache, single launch); "
                                  "static file, not measured in this run",
                "peak_source": "DMMA.8x8x4 / DFMA chains measured in this run through "
                               "dab_measure_fp64_peak (dfma %.2f, dmma %.2f TFLOP/s)"
                               % (peaks["dfma_tflops"], peaks["dmma_tflops"]),
                "flops_per_launch": aux["alg"]["forecast"]["flops"],
                "hbm_peak_gbs": hbm[0], "hbm_peak_source": hbm[1]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cps_cpu, done = time_oracle(w, 2, 0, budget_s=args.cpu_budget)
        cpu = {"value": cps_cpu, "unit": "cycles/s", "cores": cpu_threads(), "kind": "port",
               "sample": "%d full cycles of the NumPy/SciPy oracle on this box's host cores "
                         "(restatement of the reference; JAX is not installable here)" % done,
               "host_cpus": os.cpu_count()}

    N, k, S = w["N"], w["k"], w["S"]
    cps = rec["cycles_per_s"]
    line = {
        "metric": "etkf_cycles_per_s", "value": cps, "unit": "cycles/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": aux["ms"] / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(w),
        "config_detail": {
            "l2": "state (%.1f MB) is L2-resident; no flush: the cycle streams through all of it "
                  "between reuses" % (N * k * 8 / 1e6),
            "experiment": "%d-cycle run from a fresh ensemble, repeated; reset copy inside the "
                          "timed region" % BLOCK,
            "parallelism": "grid-sharded x%d" % world if world > 1 else "single GPU",
            "forecast_shape": rec["forecast_shape"]},
        "member_steps_per_s": cps * k * S,
        "point_steps_per_s": cps * k * S * N,
        "all_finite": rec["all_finite"],
        "roofline": roofline, "kernels": rec["kernels"],
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": aux["launches"],
        "clocks": clocks,
    }
    for key in ("single_gpu_cycles_per_s", "efficiency_vs_1gpu", "parity_vs_1gpu",
                "parity_vs_1gpu_after_%d_cycles" % BLOCK, "parity_note"):
        if key in rec:
            line[key] = rec[key]
    line.update(extra)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=45.0, help="seconds of CPU oracle time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 / C2 / C1 / API sub-records")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # line when the box sets NCCL_DEBUG): point fd 1 at stderr for the run and give the real stdout
    # only to the final print.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
