"""Real multi-GPU check (run under torch.distributed.run, one rank per GPU):
the sharded cycler (NCCL all-reduce + CUDA-IPC halo loads) against a single-GPU run on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dataassimbench_b200 import _native  # noqa: E402
from dataassimbench_b200.sharded import ShardedCycler  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    N, k, n_cycles = 65536, 32, 6
    bench.WORKLOADS["mg"] = dict(N=N, k=k, obs_frac=0.3, S=20, dt=0.01, sd=1.0, rho=1.0, label="mg check")
    w = bench.make_workload("mg")
    H = _native.Handle(local)
    cyc = ShardedCycler(H, w, n_cycles, rank, world)
    n_loc = N // world
    out = torch.empty((n_cycles, k, n_loc), dtype=torch.float64, device="cuda")
    cyc.run(0, n_cycles, out)
    H.sync()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    ok = True
    if rank == 0:
        full = torch.cat(gathered, dim=2).cpu().numpy()
        H1 = _native.Handle(local)
        single = ShardedCycler(H1, w, n_cycles, 0, 1)
        ref = torch.empty((n_cycles, k, N), dtype=torch.float64, device="cuda")
        single.run(0, n_cycles, ref)
        H1.sync()
        ref = ref.cpu().numpy()
        err = float(np.max(np.abs(full - ref)) / np.max(np.abs(ref)))
        ok = bool(np.isfinite(full).all()) and err < 1e-10
        print("multigpu_check world=%d device_exchange=%s rel_err=%.3e finite=%s %s" % (world, cyc.device_exchange, err, np.isfinite(full).all(),
                                                                    "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
