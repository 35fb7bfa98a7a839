"""Grid-sharded cycling on ONE GPU: `world` handles act as the ranks, each with its own stream on
the same device, and peer pointers are registered directly.  This exercises everything the
multi-GPU path adds: observation partition by owner, the K5 halo exchange through peer pointers,
the extended-layout forecast on a slab, and both ways of summing the ranks' statistics:
`exchange="device"` -- the reduction kernels store into every rank's exchange buffer and a small
kernel waits for all ranks' flags and adds the slots (what real multi-GPU runs do);
`exchange="host"` -- the caller sums `dab_cycler_stats_ptr` between stats and update (the NCCL
all-reduce's stand-in)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dataassimbench_b200 import _native
from test_gpu_kernels import cycler_case, run_gpu_cycler, rel_err
from oracle import windows as o_win


def run_emulated_ranks(case, N, k, n_cycles, sd, rho, stationary, world, exchange="host", host_out=False):
    S, dt = case['S'], case['dt']
    times = o_win.get_all_times(case['start'], case['window'], n_cycles)
    padded = np.ascontiguousarray(o_win.pad_time_indices(o_win.get_obs_indices(times, case['obs_times'], case['window'])))
    vals, sysidx = np.ascontiguousarray(case['vals']), np.ascontiguousarray(case['sysidx'])
    n_loc = N // world
    hs = []
    for r in range(world):
        h = _native.Handle(0)
        cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=rho,
                                  obs_error_sd=sd, n_obs_times=vals.shape[0], n_obs=vals.shape[1],
                                  stationary=int(stationary), n_cycles=n_cycles, t_pad=padded.shape[1],
                                  rank=r, world=world)
        h.cycler_setup(cfg, vals, sysidx, padded)
        h.cycler_set_state(np.ascontiguousarray(case['X0'][:, r * n_loc:(r + 1) * n_loc]), True)
        hs.append(h)
    for r in range(world):
        for p in range(world):
            if p != r:
                for which in ((0, 1, 2) if exchange == "device" else (0, 1)):
                    hs[r].cycler_set_peer(p, which, hs[p].cycler_buffer_ptr(which))
    assert all(h.cycler_stats_exchange() == (exchange == "device") for h in hs)

    class Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    stats = [torch.as_tensor(Dev(h.cycler_stats_ptr(), k * k + k), device='cuda') for h in hs]
    out = torch.empty((n_cycles, k, N), dtype=torch.float64, device='cuda')
    slab = [torch.empty((k, n_loc), dtype=torch.float64, device='cuda') for _ in range(world)]
    slab_host = [torch.empty((k, n_loc), dtype=torch.float64).pin_memory() for _ in range(world)]
    for c in range(n_cycles):
        for h in hs:
            h.cycler_stats(c)
        if exchange == "host":
            for h in hs:
                h.sync()
            total = torch.stack(stats).sum(dim=0)             # the all-reduce
            for s in stats:
                s.copy_(total)
            torch.cuda.synchronize()
        for r, h in enumerate(hs):
            if host_out:
                h.cycler_update_host(c, slab_host[r])      # D2H on the handle's copy stream
            else:
                h.cycler_update(c, slab[r])
        for r, h in enumerate(hs):
            h.sync()
            out[c, :, r * n_loc:(r + 1) * n_loc] = slab_host[r].cuda() if host_out else slab[r]
    final = np.empty((k, N))
    for r, h in enumerate(hs):
        part = np.empty((k, n_loc))
        h.cycler_get_state(part, True)
        final[:, r * n_loc:(r + 1) * n_loc] = part
        h.close()
    return out.cpu().numpy(), final


@pytest.mark.parametrize("exchange", ["device", "host"])
@pytest.mark.parametrize("N,k,n_obs,n_cycles,S,stationary,world", [
    (4096, 16, 1200, 5, 20, True, 2),
    (4096, 16, 1200, 5, 20, False, 4),
    (2048, 64, 600, 4, 20, True, 8),
    (640, 8, 200, 4, 10, True, 8),           # slab (80) shorter than the left halo (80): multi-owner halos
    (96, 6, 40, 4, 5, True, 4),              # slab (24) shorter than the halo (40/20): halos wrap past neighbours
])
def test_sharded_equals_single_gpu(N, k, n_obs, n_cycles, S, stationary, world, exchange):
    sd, rho = 1.0, 1.02
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, rho, stationary, seed=N + world)
    H1 = _native.Handle(0)
    ref, ref_final, _ = run_gpu_cycler(H1, case, N, k, n_cycles, sd, rho, stationary)
    H1.close()
    out, final = run_emulated_ranks(case, N, k, n_cycles, sd, rho, stationary, world, exchange)
    assert np.all(np.isfinite(out))
    # same kernels, only the summation order of the k*k+k statistics differs between 1 and G ranks
    assert rel_err(out, ref) < 1e-10
    assert rel_err(final, ref_final) < 1e-9


def test_sharded_update_host_returns_the_same_slabs():
    """dab_cycler_update_host: the per-rank slab of every analysis copied to pinned host memory on the
    copy stream gives the same bits as the device output of dab_cycler_update."""
    N, k, n_obs, n_cycles, S, world = 4096, 16, 1200, 4, 20, 2
    case = cycler_case(N, k, n_obs, n_cycles, S, 1.0, 1.02, True, seed=5)
    out_d, fin_d = run_emulated_ranks(case, N, k, n_cycles, 1.0, 1.02, True, world, "device")
    out_h, fin_h = run_emulated_ranks(case, N, k, n_cycles, 1.0, 1.02, True, world, "device", host_out=True)
    assert np.array_equal(out_d, out_h) and np.array_equal(fin_d, fin_h)
