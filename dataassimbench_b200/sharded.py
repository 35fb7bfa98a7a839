"""Grid-sharded ETKF cycling: one process (rank) per GPU.

The periodic Lorenz-96 grid is cut into `world` contiguous slabs.  Per cycle a
rank runs
    K3  gather + contraction over the observations it owns        -> partial (C, g)
        ONE exchange of k*k+k doubles: the reduction kernel stores them into every rank's exchange
        buffer over NVLink and a small kernel adds the slots in rank order (device-side, no host
        involvement; DAB_STATS_EXCHANGE=nccl selects an NCCL all-reduce via torch.distributed instead)
    K4  transform matrix (replicated, bit-identical on every rank)
    K5  transform of its slab AND of the halo columns it needs, the latter read from the
        neighbours' background through CUDA-IPC peer pointers (NVLink P2P loads)
    K1  RK4 window on slab + halo, no communication.
The exchange is the only synchronisation point: it also orders the peer reads of K5 after the
neighbours' forecasts (two alternating background buffers make that race-free).

This module holds the host logic only (partitioning, plumbing); it is testable on CPU with gloo
through `ShardPlan`.
"""
import numpy as np

from . import _native


class ShardPlan:
    """Pure-integer description of the decomposition (no device needed)."""

    def __init__(self, system_dim, world):
        if system_dim % world != 0:
            raise ValueError("system_dim %d is not divisible by world size %d" % (system_dim, world))
        self.system_dim = int(system_dim)
        self.world = int(world)
        self.n_loc = self.system_dim // self.world

    def slab(self, rank):
        return rank * self.n_loc, (rank + 1) * self.n_loc

    def owner(self, system_index):
        return np.asarray(system_index, dtype=np.int64) // self.n_loc

    def local_obs(self, rank, system_index_row, values_row=None):
        """Observations of one time owned by `rank`, order preserved: (local_idx, values, positions)."""
        idx = np.asarray(system_index_row, dtype=np.int64)
        pos = np.nonzero(self.owner(idx) == rank)[0]
        vals = None if values_row is None else np.asarray(values_row)[pos]
        return idx[pos] - rank * self.n_loc, vals, pos

    def halo_sources(self, rank, halo_l, halo_r):
        """For the extended columns e in [-halo_l, n_loc + halo_r): (owner rank, local column)."""
        n0 = rank * self.n_loc
        e = np.arange(-halo_l, self.n_loc + halo_r)
        gp = np.mod(n0 + e, self.system_dim)
        return gp // self.n_loc, gp % self.n_loc


class _DevArray:
    """Minimal __cuda_array_interface__ view of library-owned device memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class ShardedCycler:
    """Drives the C-ABI cycler on this rank; `w` is a workload dict (see bench.py / tests)."""

    def __init__(self, handle, w, n_cycles, rank=0, world=1):
        import torch
        self.torch = torch
        self.H = handle
        self.w = w
        self.rank, self.world = rank, world
        self.plan = ShardPlan(w["N"], world)
        self.k = w["k"]
        self.n_cycles = n_cycles
        cfg = self.config(n_cycles)
        self.H.cycler_setup(cfg, w["vals"], w["sysidx"], w["padded"], w.get("sd_slots"))
        lo, hi = self.plan.slab(rank)
        self.x0_dev = torch.from_numpy(np.ascontiguousarray(w["X0"][:, lo:hi])).cuda()
        self.stream = torch.cuda.ExternalStream(self.H.cycler_stream())
        self.stats = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            import os
            n_bufs = 2 if os.environ.get("DAB_STATS_EXCHANGE", "p2p") == "nccl" else 3
            handles = [self.H.cycler_ipc_export(w_) for w_ in range(n_bufs)]
            gathered = [None] * world
            dist.all_gather_object(gathered, handles)
            for r in range(world):
                if r != rank:
                    for w_ in range(n_bufs):
                        self.H.cycler_ipc_import(r, w_, gathered[r][w_])
            self.stats = torch.as_tensor(_DevArray(self.H.cycler_stats_ptr(), self.k * self.k + self.k),
                                         device="cuda")
            dist.barrier()
        # statistics summed on the device (P2P all-gather fused into the reduction kernel) unless
        # DAB_STATS_EXCHANGE=nccl asks for the torch.distributed all-reduce
        self.device_exchange = world > 1 and self.H.cycler_stats_exchange()
        self.reset()

    def config(self, n_cycles):
        w = self.w
        return _native.CycleConfig(
            system_dim=w["N"], ensemble_dim=w["k"], n_steps=w["S"], model_dt=w["dt"], forcing=w.get("forcing", 8.0),
            rho=w["rho"], obs_error_sd=w["sd"], n_obs_times=w["vals"].shape[0], n_obs=w["vals"].shape[1],
            stationary=int(w.get("stationary", 1)), n_cycles=n_cycles, t_pad=w["padded"].shape[1], rank=self.rank, world=self.world)

    def reset(self):
        """Fresh ensemble (device-to-device copy on the cycler's stream)."""
        self.H.cycler_set_state(self.x0_dev, False)

    def run(self, c0, c1, out_analysis_dev=None):
        if self.world == 1 and out_analysis_dev is None:
            self.H.cycler_run(c0, c1)
            return
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for c in range(c0, c1):
                self.H.cycler_stats(c)
                if self.world > 1 and not self.device_exchange:
                    self.dist.all_reduce(self.stats)
                out = None if out_analysis_dev is None else out_analysis_dev[c - c0]
                self.H.cycler_update(c, out)

    def state(self):
        out = self.torch.empty((self.k, self.plan.n_loc), dtype=self.torch.float64, device="cuda")
        self.H.cycler_get_state(out, False)
        self.H.sync()
        return out

    def state_is_finite(self):
        return bool(self.torch.isfinite(self.state()).all())
