"""ETKF cycler with the reference's constructor and `cycle()` signature.

Mirrors `dabench.dacycler.ETKF` (dabench/dacycler/_etkf.py:19-62) and `DACycler.cycle`
(dabench/dacycler/_dacycler.py:186-290).  The accelerated contract: the forecast
model wraps one of the package's generators -- `Lorenz96Model` runs the device-resident cycler, `Lorenz63Model` /
`SQGTurbModel` (or any Model with a `_forecast_members_device`) alternate the analysis kernels and the model's
ensemble forecast on the device -- H is the default index gather and R diagonal.  Anything else (user-supplied
dense H/B, callable h, models without a device forecast) raises `NotImplementedError` -- there is deliberately no
CPU fallback.
"""
import numpy as np

from .. import _xr, _native
from ._dacycler import DACycler, IndexOperator, ScaledIdentity


class ETKF(DACycler):
    _in_4d = False
    _uses_ensemble = True
    traj_block_bytes = 4 << 30      # device memory for return_forecast trajectories, per block of cycles

    def __init__(self, system_dim, delta_t, model_obj, B=None, R=None, H=None, h=None,
                 ensemble_dim=4, multiplicative_inflation=1.0, distributed=None):
        self.ensemble_dim = ensemble_dim
        self.multiplicative_inflation = multiplicative_inflation
        # grid sharding over the ranks of torch.distributed's default group (one process per GPU): None =
        # whenever such a group with more than one rank exists and the grid divides over it, False = never,
        # True = required (ValueError otherwise).  Not an argument of the reference (it has no multi-GPU path).
        self.distributed = distributed
        super().__init__(system_dim=system_dim, delta_t=delta_t, model_obj=model_obj,
                         B=B, R=R, H=H, h=h)

    # ------------------------------------------------------------------ the reference's overridable hooks
    # Same names, argument meaning and array orientation as dabench/dacycler/_etkf.py:64-213 (Xb is
    # (system_dim, ensemble_dim) there).  Each is one call into the C ABI; `cycle()` itself does not go
    # through them (it runs the device-resident loop), exactly as the reference's `cycle()` runs a traced
    # scan: they exist for user code that calls or composes the pieces.
    def _step_forecast(self, Xa, n_steps=1):
        """`model_obj.forecast` for every member (dabench/dacycler/_etkf.py:64-80): (last states
        `x[ensemble, index]`, trajectories `x[ensemble, time, index]` of `n_steps` points, the first being
        Xa itself) -- one `dab_l96_forecast_f64` launch for the whole ensemble."""
        import torch
        from .._runtime import handle
        gen = self._lorenz96_generator()
        if gen is None:
            raise NotImplementedError('_step_forecast needs model_obj to wrap a data.Lorenz96 generator')
        X = np.ascontiguousarray(_xr.var_array(Xa, 'x'), dtype=np.float64)
        k, N = X.shape
        dim = [d for d in _xr.var_dims(Xa, 'x') if d != 'ensemble'][0]
        steps = int(n_steps) - 1
        xin = torch.from_numpy(X).cuda()
        xout = torch.empty_like(xin)
        traj = torch.empty((max(steps, 1), k, N), dtype=torch.float64, device='cuda')
        if steps > 0:
            handle().l96_forecast(xin, xout, traj, N, k, steps, float(gen.delta_t), float(gen.forcing_term),
                                  torch.cuda.current_stream().cuda_stream)
            full = np.concatenate([X[None], traj.cpu().numpy()], axis=0)
        else:
            full = X[None]
        times = np.arange(full.shape[0]) * float(gen.delta_t)
        last = _xr.new_dataset({'x': (('ensemble', dim), full[-1].copy())})
        trajs = _xr.new_dataset({'x': (('ensemble', 'time', dim), np.ascontiguousarray(full.transpose(1, 0, 2)))},
                                coords={'time': times})
        return last, trajs

    @staticmethod
    def _index_operator(H, h, system_dim):
        if h is not None and H is None:
            raise ValueError('Only linear obs operators (H) are supported right now.')
        if isinstance(H, IndexOperator):
            return H
        H = np.asarray(H)
        if H.ndim == 1 and np.issubdtype(H.dtype, np.integer):
            return IndexOperator(H, system_dim)
        return IndexOperator.from_dense(H)

    def _apply_obsop(self, Xb, H, h):
        """`Yb = H @ Xb` (dabench/dacycler/_etkf.py:82-92) for an index-gather H: the gather kernel,
        bit-exact.  Xb is (system_dim, ensemble_dim); returns (n_obs, ensemble_dim)."""
        from ..obsop import ObsOp
        Xb = np.asarray(Xb, dtype=np.float64)
        op = self._index_operator(H, h, Xb.shape[0])
        return ObsOp.gather(np.ascontiguousarray(Xb.T), op.indices, op.mask).T

    def _compute_analysis(self, Xb, Y, H, h, R, rho=1.0):
        """The ETKF analysis (dabench/dacycler/_etkf.py:94-163) for an index-gather H and R = sd**2 I:
        `dab_etkf_analysis_f64` (gather + DMMA contraction, k x k transform matrix, DMMA transform).
        Xb and the result are (system_dim, ensemble_dim)."""
        import torch
        from .._runtime import handle
        Xb = np.asarray(Xb, dtype=np.float64)
        N, k = Xb.shape
        y = np.ascontiguousarray(np.ravel(Y), dtype=np.float64)
        empty_R = len(R) == 0                                   # the reference's zero branch (:149-152)
        n_y = 0 if empty_R else y.shape[0]
        Xd = torch.from_numpy(np.ascontiguousarray(Xb.T)).cuda()
        Xa = torch.empty_like(Xd)
        yd = idxd = md = None
        sd = 1.0
        if n_y > 0:
            op = self._index_operator(H, h, N)
            if op.indices.shape[0] != n_y:
                raise ValueError('H has %d rows for %d observations' % (op.indices.shape[0], n_y))
            sd = ScaledIdentity.sd_of(R)
            if np.ndim(sd) != 0:                               # diagonal R with unequal entries
                if sd.shape[0] != n_y:
                    raise ValueError('R has %d rows for %d observations' % (sd.shape[0], n_y))
                sd = torch.from_numpy(np.ascontiguousarray(sd)).cuda()
            yd = torch.from_numpy(y).cuda()
            idxd = torch.from_numpy(op.indices).cuda()
            md = None if op.mask is None else torch.from_numpy(op.mask.astype(np.uint8)).cuda()
        handle().etkf_analysis(Xd, Xa, yd, idxd, md, sd if torch.is_tensor(sd) else float(sd), float(rho), N, k, n_y,
                               torch.cuda.current_stream().cuda_stream)
        return np.ascontiguousarray(Xa.cpu().numpy().T)

    def _cycle_obsop(self, Xb_ds, obs_values, obs_loc_indices, obs_time_mask, obs_loc_mask,
                     H=None, h=None, R=None, B=None):
        """One analysis on a Dataset (dabench/dacycler/_etkf.py:165-213): default operators resolved as
        the reference does, masks applied to H, `_compute_analysis`, result written back as
        `x[ensemble, i]`."""
        if H is None and h is None:
            if self.H is None:
                if self.h is None:
                    H = self._calc_default_H(obs_values, obs_loc_indices)
                else:
                    h = self.h
            else:
                H = self.H
        if R is None:
            R = self._calc_default_R(obs_values, self.obs_error_sd) if self.R is None else self.R
        Xb = np.asarray(_xr.var_array(Xb_ds, 'x'), dtype=np.float64).T
        n_sys, n_ens = Xb.shape
        assert n_ens == self.ensemble_dim, (
            'cycle:: model_forecast must have dimension {}x{}').format(self.ensemble_dim, self.system_dim)
        op = self._index_operator(H, h, n_sys)
        op = op.masked(np.ravel(obs_time_mask)).masked(np.ravel(obs_loc_mask))
        Xa = self._compute_analysis(Xb=Xb, Y=obs_values, H=op, h=None, R=R, rho=self.multiplicative_inflation)
        return _xr.new_dataset({'x': (('ensemble', 'i'), np.ascontiguousarray(Xa.T))})

    # ------------------------------------------------------------------ marshalling (CPU-testable)
    def _plan(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd=None,
              analysis_window=0.2, analysis_time_in_window=None):
        """Everything `dab_cycler_setup` needs, as plain arrays."""
        if self.h is not None and self.H is None:
            # dabench/dacycler/_dacycler.py:103-104
            raise ValueError('Only linear obs operators (H) are supported right now.')
        if self.H is not None or self.B is not None:
            raise NotImplementedError('user-supplied H / B matrices are outside the accelerated path '
                                      '(default index-gather H only)')
        gen = self._lorenz96_generator()
        generic = gen is None
        if generic and not callable(getattr(self.model_obj, '_forecast_members_device', None)):
            raise NotImplementedError('the accelerated ETKF.cycle() needs model_obj to be a Model wrapping a '
                                      'dataassimbench_b200.data generator (model.Lorenz96Model, Lorenz63Model, '
                                      'SQGTurbModel), i.e. one with a device forecast for the whole ensemble')
        w = self._prepare_cycle(input_state, start_time, obs_vector, n_cycles, obs_error_sd,
                                analysis_window, analysis_time_in_window)
        if self._data_vars != ['x'] or self._observed_vars != ['x']:
            raise NotImplementedError("single-variable states named 'x' only (as the reference's ETKF, "
                                      "dabench/dacycler/_etkf.py:213)")
        # R: the default obs_error_sd**2 * I, a vector obs_error_sd (one sigma per observation slot: a diagonal R),
        # or a user R that is diagonal (dabench/dacycler/_etkf.py:142 takes pinv(R): for a diagonal R that is
        # 1 / sd**2 entry by entry, 0 where sd == 0)
        sd = ScaledIdentity.sd_of(self.R) if self.R is not None else self.obs_error_sd
        sd_slots = None
        if np.ndim(sd) != 0:
            sd = np.asarray(sd, dtype=np.float64)
            n_t = len(w['obs_times'])
            raw_x = _xr.var_data(obs_vector, 'x')
            n_o = int(np.prod(tuple(raw_x.shape))) // max(n_t, 1)
            if sd.shape not in ((n_o,), (n_t, n_o)):
                raise ValueError('a vector obs_error_sd / diagonal R needs one entry per observation slot: '
                                 'shape (%d,) or (%d, %d), got %s' % (n_o, n_t, n_o, sd.shape))
            sd_slots = np.ascontiguousarray(np.broadcast_to(sd, (n_t, n_o)))
            sd = 1.0
        X0 = np.ascontiguousarray(_xr.var_array(input_state, 'x'), dtype=np.float64)
        if X0.ndim != 2:
            raise ValueError("input_state['x'] must have dims ('ensemble', <state>)")
        n_ens, n_sys = X0.shape
        # dabench/dacycler/_etkf.py:196-199
        assert n_ens == self.ensemble_dim, (
            'cycle:: model_forecast must have dimension {}x{}').format(self.ensemble_dim, self.system_dim)
        raw = _xr.var_data(obs_vector, 'x')
        if getattr(raw, 'is_cuda', False):
            # observations sampled on the device (Observer(..., keep_on_device=True)): they stay there
            vals = raw.reshape(len(w['obs_times']), -1).contiguous()
        else:
            vals = np.asarray(_xr.var_array(obs_vector, 'x'), dtype=np.float64)
            vals = np.ascontiguousarray(vals.reshape(len(w['obs_times']), -1))
        sysidx = np.asarray(_xr.var_array(obs_vector, 'system_index'))
        sysidx = np.ascontiguousarray(sysidx.reshape(vals.shape), dtype=np.int64)
        cfg = _native.CycleConfig(
            system_dim=n_sys, ensemble_dim=n_ens, n_steps=self.steps_per_window - 1,
            model_dt=float(gen.delta_t) if gen is not None else float(self.delta_t),
            forcing=float(gen.forcing_term) if gen is not None else 0.0,
            rho=float(self.multiplicative_inflation), obs_error_sd=float(sd),
            n_obs_times=vals.shape[0], n_obs=vals.shape[1], stationary=int(w['stationary']),
            n_cycles=int(n_cycles), t_pad=w['padded'].shape[1], rank=0, world=1)
        state_dim = [d for d in _xr.var_dims(input_state, 'x') if d != 'ensemble']
        return dict(cfg=cfg, X0=X0, vals=vals, sysidx=sysidx, sd_slots=sd_slots, generic=generic,
                    stationary=bool(w['stationary']), padded=np.ascontiguousarray(w['padded']), state_dim=state_dim[0] if state_dim else 'index')

    # ------------------------------------------------------------------ grid sharding (one rank per GPU)
    def _shard_group(self, N, S):
        """(rank, world) of the run's process group when this call is to be grid-sharded, else None."""
        if self.distributed is False:
            return None
        try:
            import torch.distributed as dist
            ready = dist.is_available() and dist.is_initialized()
        except Exception:
            ready = False
        world = dist.get_world_size() if ready else 1
        # every rank's slab must hold the halo its neighbours read from it (8 S points to the left)
        ok = world > 1 and world <= 8 and N % world == 0 and N // world >= max(12 * S, 32)
        if not ok:
            if self.distributed is True:
                raise ValueError('distributed=True needs an initialised torch.distributed group of 2..8 ranks and a '
                                 'grid that divides over them into slabs of at least 12 steps_per_window points')
            return None
        return dist.get_rank(), world

    def _cycle_sharded(self, p, group):
        """The same experiment with the grid cut into one slab per rank (`sharded.ShardedCycler`):
        every rank passes the SAME inputs, runs the cycles on its slab -- the ranks exchange k*k+k statistics per
        cycle and read each other's halo columns over NVLink -- and gets the full analyses back (one all-gather
        of the slabs per block of cycles)."""
        import torch
        import torch.distributed as dist
        from .._runtime import handle
        from ..sharded import ShardedCycler
        rank, world = group
        cfg = p['cfg']
        k, N, n_cycles = cfg.ensemble_dim, cfg.system_dim, cfg.n_cycles
        w = dict(N=N, k=k, S=cfg.n_steps, dt=cfg.model_dt, rho=cfg.rho, sd=cfg.obs_error_sd, X0=p['X0'],
                 vals=p['vals'], sysidx=p['sysidx'], padded=p['padded'], forcing=cfg.forcing,
                 stationary=cfg.stationary, sd_slots=p['sd_slots'])
        cyc = ShardedCycler(handle(), w, n_cycles, rank, world)
        n_loc = N // world
        out = np.empty((n_cycles, k, N))
        blk = int(max(1, min(n_cycles, self.traj_block_bytes // (k * N * 8))))
        slab = torch.empty((blk, k, n_loc), dtype=torch.float64, device='cuda')
        full = torch.empty((world, blk, k, n_loc), dtype=torch.float64, device='cuda')
        for c0 in range(0, n_cycles, blk):
            c1 = min(n_cycles, c0 + blk)
            cyc.run(c0, c1, slab[:c1 - c0])
            handle().sync()
            dist.all_gather_into_tensor(full, slab)
            got = full[:, :c1 - c0].permute(1, 2, 0, 3).reshape(c1 - c0, k, N)
            out[c0:c1] = got.cpu().numpy()
        return out

    # ------------------------------------------------------------------ any other model with a device forecast
    def _cycle_generic(self, p, return_forecast):
        """The reference's per-cycle body (dabench/dacycler/_dacycler.py:110-141) for a model other than Lorenz-96:
        the observation rows of every cycle are laid out once (time mask = rows dropped, NaN slots of a moving
        network masked), then analysis (`dab_etkf_analysis_f64`) and the model's ensemble forecast alternate on the
        device; the ensemble never visits the host.  What a cycle returns is row 0 of the model's trajectory, as
        in the reference (:288-290): the analysis itself for the package's wrappers."""
        import torch
        from .._runtime import handle
        cfg = p['cfg']
        k, N, n_cycles = cfg.ensemble_dim, cfg.system_dim, cfg.n_cycles
        vals = p['vals'].cpu().numpy() if getattr(p['vals'], 'is_cuda', False) else p['vals']
        H = handle()
        stream = torch.cuda.current_stream().cuda_stream
        rows_y, rows_i, rows_s, offs = [], [], [], [0]
        for c in range(n_cycles):
            t = p['padded'][c]
            t = t[t > 0] - 1
            y, idx = vals[t].ravel(), p['sysidx'][t].ravel()
            sds = p['sd_slots'][t].ravel() if p['sd_slots'] is not None else None
            if not p['stationary']:
                keep = ~np.isnan(y)
                y, idx = y[keep], idx[keep]
                sds = sds[keep] if sds is not None else None
            if sds is not None:
                rows_s.append(sds)
            if idx.size and (idx.min() < 0 or idx.max() >= N):
                raise ValueError('system_index out of range for system_dim %d' % N)
            rows_y.append(y); rows_i.append(idx); offs.append(offs[-1] + y.size)
        y_all = torch.from_numpy(np.concatenate(rows_y) if rows_y else np.zeros(0)).cuda()
        i_all = torch.from_numpy((np.concatenate(rows_i) if rows_i else np.zeros(0)).astype(np.int64)).cuda()
        # a diagonal R with unequal entries: one sigma per row, on the device (dab_etkf_analysis_diag_f64)
        s_all = torch.from_numpy(np.ascontiguousarray(np.concatenate(rows_s), dtype=np.float64)).cuda() if rows_s else None
        X = torch.from_numpy(p['X0']).cuda()
        Xa = torch.empty_like(X)
        out = None
        sd, rho = float(cfg.obs_error_sd), float(cfg.rho)
        # a window without a valid observation is NOT the reference's empty-R branch (its R still has the padded
        # rows, all masked: C = 0, inflation only): one masked dummy row says so to the analysis entry point
        y0 = torch.zeros(1, dtype=torch.float64, device='cuda')
        i0 = torch.zeros(1, dtype=torch.int64, device='cuda')
        m0 = torch.zeros(1, dtype=torch.uint8, device='cuda')
        empty_R = vals.shape[1] == 0 or p['padded'].shape[1] == 0
        for c in range(n_cycles):
            a, b = offs[c], offs[c + 1]
            if b > a:
                H.etkf_analysis(X, Xa, y_all[a:b], i_all[a:b], None, sd if s_all is None else s_all[a:b], rho, N, k,
                                b - a, stream)
            elif empty_R:
                H.etkf_analysis(X, Xa, None, None, None, sd, rho, N, k, 0, stream)
            else:
                H.etkf_analysis(X, Xa, y0, i0, m0, sd, rho, N, k, 1, stream)
            X, first, rows = self.model_obj._forecast_members_device(Xa, self.steps_per_window, return_forecast)
            # what the reference keeps of a cycle's trajectory: all but its last row (:281-287), or row 0 (:288-290)
            keep = rows[:-1].permute(1, 0, 2) if return_forecast else first
            if out is None:
                out = torch.empty((n_cycles,) + tuple(keep.shape), dtype=torch.float64, device='cuda')
            out[c] = keep
            X = X.contiguous()
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy()

    # ------------------------------------------------------------------ the call
    def cycle(self, input_state, start_time, obs_vector, n_cycles, obs_error_sd=None,
              analysis_window=0.2, analysis_time_in_window=None, return_forecast=False):
        """Analysis + forecast, `n_cycles` times (dabench/dacycler/_dacycler.py:186-290).

        Returns a Dataset `x[cycle, ensemble, index]` (analyses) or, with return_forecast=True,
        `x[cycle, ensemble, cycle_timestep, index]` where cycle_timestep 0 is the analysis and the
        last forecast point (the next background) is dropped, as in the reference (:281-290)."""
        import torch
        from .._runtime import handle
        p = self._plan(input_state, start_time, obs_vector, n_cycles, obs_error_sd,
                       analysis_window, analysis_time_in_window)
        cfg = p['cfg']
        k, N, S = cfg.ensemble_dim, cfg.system_dim, cfg.n_steps
        if p['generic']:
            dims = ('cycle', 'ensemble', 'cycle_timestep', p['state_dim']) if return_forecast else \
                ('cycle', 'ensemble', p['state_dim'])
            return _xr.new_dataset({'x': (dims, self._cycle_generic(p, return_forecast))})
        group = self._shard_group(N, S)
        if group is not None:
            if return_forecast:
                raise NotImplementedError('return_forecast=True is single-GPU only')
            return _xr.new_dataset({'x': (('cycle', 'ensemble', p['state_dim']), self._cycle_sharded(p, group))})
        H = handle()
        # pinned directly (no pageable detour); torch's caching host allocator recycles the block of a result the
        # caller has dropped, so repeated calls do not pay the page-locking again
        out = torch.empty((n_cycles, k, N), dtype=torch.float64, pin_memory=True)
        dim = p['state_dim']
        if not return_forecast:
            if getattr(p['vals'], 'is_cuda', False):
                H.cycler_setup(cfg, p['vals'], p['sysidx'], p['padded'], p['sd_slots'])     # dab_cycler_setup_dev
                H.cycler_set_state(p['X0'], True)
                H.cycler_run(0, n_cycles, out)
                H.sync()
            else:
                H.etkf_cycle_host(cfg, p['X0'], p['vals'], p['sysidx'], p['padded'], out, p['sd_slots'])
            return _xr.new_dataset({'x': (('cycle', 'ensemble', dim), out.numpy())})
        # The trajectories are S*k*N*8 bytes per cycle (671 MB at N=65536, k=64): they are produced in
        # blocks of cycles that fit `traj_block_bytes` of device memory and streamed to the host block
        # by block, instead of materialising n_cycles of them in HBM (the reference builds the whole
        # array inside its scan, dabench/dacycler/_dacycler.py:276-290).
        H.cycler_setup(cfg, p['vals'], p['sysidx'], p['padded'], p['sd_slots'])
        H.cycler_set_state(p['X0'], True)
        full = np.empty((n_cycles, k, max(S, 1), N))
        per_cycle = max(S, 1) * k * N * 8
        blk = int(max(1, min(n_cycles, self.traj_block_bytes // per_cycle)))
        traj = torch.empty((blk, max(S, 1), k, N), dtype=torch.float64, device='cuda') if S > 0 else None
        for c0 in range(0, n_cycles, blk):
            c1 = min(n_cycles, c0 + blk)
            H.cycler_run(c0, c1, out[c0:c1], traj[:c1 - c0] if S > 0 else None)
            H.sync()
            if S > 1:
                full[c0:c1, :, 1:, :] = traj[:c1 - c0, :S - 1].permute(0, 2, 1, 3).cpu().numpy()
        full[:, :, 0, :] = out.numpy()
        return _xr.new_dataset({'x': (('cycle', 'ensemble', 'cycle_timestep', dim), full)})
