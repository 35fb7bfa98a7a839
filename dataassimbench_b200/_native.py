"""ctypes binding of the C ABI in include/dab_b200.h (libdab_b200.so).

There is no CPU fallback: if the CUDA library is missing or no B200 is present, loading or
`Handle()` raises.  torch is used only to own device memory and streams (`tensor.data_ptr()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdab_b200.so")

IPC_HANDLE_BYTES = 64
MAX_ENSEMBLE = 128
MAX_RANKS = 8


class NativeLibraryMissing(ImportError):
    pass


class DabError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libdab_b200 error %d: %s" % (code, msg))
        self.code = code


class CycleConfig(C.Structure):
    """struct dab_cycle_config (include/dab_b200.h)."""
    _fields_ = [
        ("system_dim", C.c_int64),
        ("ensemble_dim", C.c_int32),
        ("n_steps", C.c_int32),
        ("model_dt", C.c_double),
        ("forcing", C.c_double),
        ("rho", C.c_double),
        ("obs_error_sd", C.c_double),
        ("n_obs_times", C.c_int64),
        ("n_obs", C.c_int64),
        ("stationary", C.c_int32),
        ("n_cycles", C.c_int32),
        ("t_pad", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_SIGNATURES = {
    # name: (restype, argtypes)
    "dab_version": (C.c_int, []),
    "dab_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "dab_destroy": (C.c_int, [_P]),
    "dab_sync": (C.c_int, [_P]),
    "dab_last_error": (C.c_char_p, [_P]),
    "dab_set_eig_method": (C.c_int, [_P, C.c_int]),
    "dab_set_pdl": (C.c_int, [C.c_int]),
    "dab_l96_forecast_f64": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, C.c_int, C.c_double, C.c_double, _P]),
    "dab_l63_forecast_f64": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "dab_sqg_setup": (C.c_int, [_P, _P]),
    "dab_sqg_forecast_c128": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "dab_sqg_forecast_grid_f64": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "dab_obs_gather_f64": (C.c_int, [_P, _P, _P, _P, _P, _I64, C.c_int, _I64, _P]),
    "dab_observe_f64": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _P, C.c_int, _P, _I64, _P, _P, _P]),
    "dab_var3d_analysis_f64": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I64, C.c_double, _P]),
    "dab_var3d_analysis_b_f64": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I64, C.c_double, _P, C.c_double, C.c_int, _P, _P]),
    "dab_etkf_stats_f64": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, _I64, C.c_int, _I64, _P, _P]),
    "dab_etkf_transform_matrix_f64": (C.c_int, [_P, _P, C.c_int, C.c_double, C.c_int, _P, _P, _P, _P]),
    "dab_debug_l96_chain_plan": (C.c_int, [_I64, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _I64, _P, _I64]),
    "dab_etkf_analysis_f64": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_double, C.c_double, _I64, C.c_int, _I64, _P]),
    "dab_etkf_analysis_diag_f64": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_double, _I64, C.c_int, _I64, _P]),
    "dab_cycler_setup": (C.c_int, [_P, C.POINTER(CycleConfig), _P, _P, _P]),
    "dab_cycler_setup_dev": (C.c_int, [_P, C.POINTER(CycleConfig), _P, _P, _P]),
    "dab_cycler_setup_diag": (C.c_int, [_P, C.POINTER(CycleConfig), _P, _P, _P, _P]),
    "dab_etkf_cycle_host_diag_f64": (C.c_int, [_P, C.POINTER(CycleConfig), _P, _P, _P, _P, _P, _P]),
    "dab_cycler_set_state": (C.c_int, [_P, _P, C.c_int]),
    "dab_cycler_get_state": (C.c_int, [_P, _P, C.c_int]),
    "dab_cycler_run": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "dab_cycler_stats": (C.c_int, [_P, C.c_int]),
    "dab_cycler_stats_exchange": (C.c_int, [_P]),
    "dab_cycler_stats_ptr": (_P, [_P]),
    "dab_cycler_update": (C.c_int, [_P, C.c_int, _P]),
    "dab_cycler_update_host": (C.c_int, [_P, C.c_int, _P]),
    "dab_cycler_set_truth": (C.c_int, [_P, _P, _P]),
    "dab_cycler_ipc_export": (C.c_int, [_P, C.c_int, _P]),
    "dab_cycler_ipc_import": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "dab_cycler_buffer_ptr": (_P, [_P, C.c_int]),
    "dab_cycler_set_peer": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "dab_cycler_stream": (_P, [_P]),
    "dab_cycler_background_ptr": (_P, [_P]),
    "dab_cycler_kernel_launches": (C.c_int, [_P]),
    "dab_cycler_forecast_shape": (C.c_int, [_P]),
    "dab_cycler_compact_contractions": (C.c_int, [_P]),
    "dab_cycler_set_profiling": (C.c_int, [_P, C.c_int]),
    "dab_cycler_read_profile": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "dab_measure_fp64_peak": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dab_etkf_cycle_host_f64": (C.c_int, [_P, C.POINTER(CycleConfig), _P, _P, _P, _P, _P]),
    "dab_etkf_cycle_batched_f64": (C.c_int, [_P, C.POINTER(CycleConfig), C.c_int, _P, _P, _P, _P, C.c_int,
                                             _P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class SqgConfig(C.Structure):
    """dab_sqg_config (include/dab_b200.h)."""
    _fields_ = [("n", C.c_int32), ("ekman", C.c_int32), ("symmetric", C.c_int32), ("reserved", C.c_int32),
                ("delta_t", C.c_double), ("inv_tdiab", C.c_double), ("r", C.c_double),
                ("kx", C.c_void_p), ("ly", C.c_void_p), ("hovermu", C.c_void_p), ("tanhmu", C.c_void_p),
                ("sinhmu", C.c_void_p), ("ksqlsq", C.c_void_p), ("hyperdiff", C.c_void_p), ("pvspec_eq", C.c_void_p)]


def load():
    """Load libdab_b200.so (built in-tree by `_build.build()`); raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: build it with `python -m dataassimbench_b200._build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(x):
    """Device/host address of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError("cannot take the address of %r" % type(x))


class Handle:
    """RAII wrapper of `dab_handle*`."""

    def __init__(self, device=0):
        self._lib = load()
        self._h = _P()
        rc = self._lib.dab_create(C.byref(self._h), int(device))
        if rc != 0:
            msg = self._lib.dab_last_error(None)
            self._h = None
            raise DabError(rc, msg.decode() if msg else "dab_create failed")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.dab_last_error(self._h)
            raise DabError(rc, msg.decode() if msg else "?")

    # ---- thin 1:1 wrappers ------------------------------------------------------------
    def sync(self):
        self._check(self._lib.dab_sync(self._h))

    def l96_forecast(self, x_in, x_out, traj, N, k, n_steps, dt, F, stream=None):
        self._check(self._lib.dab_l96_forecast_f64(self._h, _ptr(x_in), _ptr(x_out), _ptr(traj),
                                                   N, k, n_steps, dt, F, _ptr(stream)))

    def l63_forecast(self, x_in, x_out, traj, k, n_steps, dt, sigma=10.0, rho=28.0, beta=8.0 / 3.0, stream=None):
        self._check(self._lib.dab_l63_forecast_f64(self._h, _ptr(x_in), _ptr(x_out), _ptr(traj), k, n_steps,
                                                   dt, sigma, rho, beta, _ptr(stream)))

    def sqg_setup(self, n, delta_t, inv_tdiab, r, ekman, symmetric, kx, ly, hovermu, tanhmu, sinhmu, ksqlsq,
                  hyperdiff, pvspec_eq):
        """Tables of the surface-QG generator (host arrays; copied to the device by the call)."""
        import numpy as np
        self._sqg_owner = None          # data.SQGTurb marks the handle after a successful set-up of ITS tables
        f8 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        arrs = [f8(kx), f8(ly), f8(hovermu), f8(tanhmu), f8(sinhmu), f8(ksqlsq), f8(hyperdiff),
                np.ascontiguousarray(np.asarray(pvspec_eq, dtype=np.complex128))]
        nk = n // 2 + 1
        assert arrs[0].shape == (nk,) and arrs[1].shape == (n,) and arrs[7].shape == (2, n, nk)
        assert all(a.shape == (n, nk) for a in arrs[2:7])
        cfg = SqgConfig(n=n, ekman=int(ekman), symmetric=int(symmetric), reserved=0, delta_t=delta_t,
                        inv_tdiab=inv_tdiab, r=r, **{name: a.ctypes.data for name, a in zip(
                            ("kx", "ly", "hovermu", "tanhmu", "sinhmu", "ksqlsq", "hyperdiff", "pvspec_eq"), arrs)})
        self._check(self._lib.dab_sqg_setup(self._h, C.byref(cfg)))

    def sqg_forecast(self, spec_in, spec_out, traj_grid, members, n_steps, stream=None):
        self._check(self._lib.dab_sqg_forecast_c128(self._h, _ptr(spec_in), _ptr(spec_out), _ptr(traj_grid),
                                                    int(members), int(n_steps), stream))

    def sqg_forecast_grid(self, grid_in, grid_out, spec_out, traj_grid, members, n_steps, stream=None):
        self._check(self._lib.dab_sqg_forecast_grid_f64(self._h, _ptr(grid_in), _ptr(grid_out), _ptr(spec_out),
                                                        _ptr(traj_grid), int(members), int(n_steps), stream))

    def obs_gather(self, X, idx, mask, Yb, N, k, n_y, stream=None):
        self._check(self._lib.dab_obs_gather_f64(self._h, _ptr(X), _ptr(idx), _ptr(mask), _ptr(Yb),
                                                 N, k, n_y, _ptr(stream)))

    def observe(self, state, N, n_state_times, time_pos, n_t, loc, stationary, count, n_obs, errors, values,
                stream=None):
        """Observer.observe sampling on the device: values = state[time_pos][loc] + errors (NaN tails)."""
        self._check(self._lib.dab_observe_f64(self._h, _ptr(state), N, n_state_times, _ptr(time_pos), n_t,
                                              _ptr(loc), int(bool(stationary)), _ptr(count), n_obs,
                                              _ptr(errors), _ptr(values), _ptr(stream)))

    def var3d_analysis(self, xb, xa, N, loc, val, n_rows, sd, stream=None):
        """3D-Var analysis for default H / R / B: xa = (xb + H^T y / sd^2) / (1 + diag(H^T H) / sd^2)."""
        self._check(self._lib.dab_var3d_analysis_f64(self._h, _ptr(xb), _ptr(xa), N, _ptr(loc), _ptr(val),
                                                     n_rows, sd, _ptr(stream)))

    def var3d_analysis_b(self, xb, xa, N, loc, val, n_rows, sd, B, tol=1e-5, maxiter=1000, stream=None):
        """3D-Var analysis with a dense user B [N][N] (device): conjugate gradients; returns the iteration count."""
        iters = C.c_int(0)
        self._check(self._lib.dab_var3d_analysis_b_f64(self._h, _ptr(xb), _ptr(xa), N, _ptr(loc), _ptr(val), n_rows,
                                                       sd, _ptr(B), tol, int(maxiter), C.byref(iters), _ptr(stream)))
        return iters.value

    def etkf_stats(self, X, y, idx, mask, sd, N, k, n_y, stats, stream=None):
        self._check(self._lib.dab_etkf_stats_f64(self._h, _ptr(X), _ptr(y), _ptr(idx), _ptr(mask),
                                                 sd, N, k, n_y, _ptr(stats), _ptr(stream)))

    EIG_AUTO, EIG_JACOBI, EIG_NEWTON_SCHULZ = 0, 1, 2

    def set_eig_method(self, method):
        """K4 solver for 32 < k <= 128: EIG_AUTO (cluster Newton-Schulz, Jacobi fallback) or EIG_JACOBI."""
        self._check(self._lib.dab_set_eig_method(self._h, int(method)))

    def set_pdl(self, on):
        """Programmatic dependent launch between the kernels of a cycle (process-wide switch)."""
        self._check(self._lib.dab_set_pdl(int(bool(on))))

    def etkf_transform_matrix(self, stats, k, rho, empty_R, T, lam=None, sweeps=None, stream=None):
        self._check(self._lib.dab_etkf_transform_matrix_f64(self._h, _ptr(stats), k, rho, int(empty_R),
                                                            _ptr(T), _ptr(lam), _ptr(sweeps), _ptr(stream)))

    def etkf_analysis(self, Xb, Xa, y, idx, mask, sd, rho, N, k, n_y, stream=None):
        """`sd`: a float (R = sd^2 I) or a device array [n_y] (diagonal R, one sigma per observation)."""
        if isinstance(sd, (int, float)):
            self._check(self._lib.dab_etkf_analysis_f64(self._h, _ptr(Xb), _ptr(Xa), _ptr(y), _ptr(idx),
                                                        _ptr(mask), sd, rho, N, k, n_y, _ptr(stream)))
        else:
            self._check(self._lib.dab_etkf_analysis_diag_f64(self._h, _ptr(Xb), _ptr(Xa), _ptr(y), _ptr(idx),
                                                             _ptr(mask), _ptr(sd), rho, N, k, n_y, _ptr(stream)))

    def cycler_setup(self, cfg, obs_values, system_index, padded_time_idx, obs_error_sd=None):
        """`obs_error_sd`: None (cfg.obs_error_sd) or a host array [n_obs_times][n_obs] (diagonal R per slot)."""
        if getattr(obs_values, "is_cuda", False):
            if obs_error_sd is not None:
                raise NotImplementedError("device-resident observations with a per-slot obs_error_sd")
            self._check(self._lib.dab_cycler_setup_dev(self._h, C.byref(cfg), _ptr(obs_values),
                                                       _ptr(system_index), _ptr(padded_time_idx)))
        elif obs_error_sd is None:
            self._check(self._lib.dab_cycler_setup(self._h, C.byref(cfg), _ptr(obs_values),
                                                   _ptr(system_index), _ptr(padded_time_idx)))
        else:
            self._check(self._lib.dab_cycler_setup_diag(self._h, C.byref(cfg), _ptr(obs_values), _ptr(system_index),
                                                        _ptr(padded_time_idx), _ptr(obs_error_sd)))

    def cycler_set_state(self, x, is_host):
        self._check(self._lib.dab_cycler_set_state(self._h, _ptr(x), int(is_host)))

    def cycler_get_state(self, x, is_host):
        self._check(self._lib.dab_cycler_get_state(self._h, _ptr(x), int(is_host)))

    def cycler_run(self, c0, c1, out_analysis=None, out_traj=None):
        self._check(self._lib.dab_cycler_run(self._h, c0, c1, _ptr(out_analysis), _ptr(out_traj)))

    def cycler_stats(self, cycle):
        self._check(self._lib.dab_cycler_stats(self._h, cycle))

    def cycler_stats_exchange(self):
        """True when dab_cycler_stats sums the ranks' statistics itself (device-side exchange over peer memory)."""
        return bool(self._lib.dab_cycler_stats_exchange(self._h))

    def cycler_stats_ptr(self):
        return self._lib.dab_cycler_stats_ptr(self._h)

    def cycler_forecast_shape(self):
        """(shape code, launches per window) the autotuner picked for the forecast kernel."""
        v = self._lib.dab_cycler_forecast_shape(self._h)
        return v // 10, v % 10

    def cycler_compact_contractions(self):
        """Contractions since setup that read the compact Yb the forecast kernels wrote (no gather)."""
        return self._lib.dab_cycler_compact_contractions(self._h)

    def cycler_update(self, cycle, out_analysis_dev=None):
        self._check(self._lib.dab_cycler_update(self._h, cycle, _ptr(out_analysis_dev)))

    def cycler_set_truth(self, truth_dev, sse_out_dev):
        """Score every analysis on the device: sse_out_dev[cycle] = sum_e (ensemble mean - truth[cycle])^2."""
        self._check(self._lib.dab_cycler_set_truth(self._h, _ptr(truth_dev), _ptr(sse_out_dev)))

    def cycler_update_host(self, cycle, out_analysis_host=None):
        """As cycler_update, the slab of the analysis going to (pinned) host memory on the copy stream."""
        self._check(self._lib.dab_cycler_update_host(self._h, cycle, _ptr(out_analysis_host)))

    def cycler_ipc_export(self, which):
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.dab_cycler_ipc_export(self._h, which, buf))
        return bytes(buf.raw)

    def cycler_ipc_import(self, peer_rank, which, handle_bytes):
        buf = C.create_string_buffer(bytes(handle_bytes), IPC_HANDLE_BYTES)
        self._check(self._lib.dab_cycler_ipc_import(self._h, peer_rank, which, buf))

    def cycler_buffer_ptr(self, which):
        return self._lib.dab_cycler_buffer_ptr(self._h, which)

    def cycler_set_peer(self, peer_rank, which, ptr):
        self._check(self._lib.dab_cycler_set_peer(self._h, peer_rank, which, _ptr(ptr)))

    def cycler_stream(self):
        return self._lib.dab_cycler_stream(self._h)

    def cycler_background_ptr(self):
        return self._lib.dab_cycler_background_ptr(self._h)

    def kernel_launches(self):
        return self._lib.dab_cycler_kernel_launches(self._h)

    PROFILE_KERNELS = ("contraction", "reduction", "eigensolver", "transform", "forecast")

    def set_profiling(self, on):
        self._check(self._lib.dab_cycler_set_profiling(self._h, int(on)))

    def read_profile(self):
        """{kernel: (total_ms, launches)} since profiling was switched on / last read."""
        ms = (C.c_double * 5)()
        cnt = (C.c_int * 5)()
        self._check(self._lib.dab_cycler_read_profile(self._h, ms, cnt))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(self.PROFILE_KERNELS)}

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.dab_measure_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dmma_tflops": b.value}

    def etkf_cycle_batched(self, cfg, batch, rho, x0, obs_values, system_index, shared_index,
                           padded_time_idx, out_analysis=None, out_mean=None, out_final=None):
        self._check(self._lib.dab_etkf_cycle_batched_f64(
            self._h, C.byref(cfg), int(batch), _ptr(rho), _ptr(x0), _ptr(obs_values), _ptr(system_index),
            int(shared_index), _ptr(padded_time_idx), _ptr(out_analysis), _ptr(out_mean), _ptr(out_final)))

    def etkf_cycle_host(self, cfg, x0, obs_values, system_index, padded_time_idx, out_analysis, obs_error_sd=None):
        if obs_error_sd is None:
            self._check(self._lib.dab_etkf_cycle_host_f64(self._h, C.byref(cfg), _ptr(x0), _ptr(obs_values),
                                                          _ptr(system_index), _ptr(padded_time_idx),
                                                          _ptr(out_analysis)))
        else:
            self._check(self._lib.dab_etkf_cycle_host_diag_f64(self._h, C.byref(cfg), _ptr(x0), _ptr(obs_values),
                                                               _ptr(system_index), _ptr(padded_time_idx),
                                                               _ptr(obs_error_sd), _ptr(out_analysis)))
