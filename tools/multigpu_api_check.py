"""Grid sharding through the public API (run under torch.distributed.run, one rank per GPU): every rank calls
`dacycler.ETKF(...).cycle()` with the same inputs; the result must equal a single-GPU call (distributed=False).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29512 tools/multigpu_api_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dataassimbench_b200 as dab  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, k, n_cycles = 8192, 16, 5      # (6 x 0.2 trips the reference's float-arange quirk, _windows.py)
    gen = dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True)
    nature = gen.generate(n_steps=n_cycles * 20 + 1)
    obs = dab.observer.Observer(nature, times=nature['time'].values[np.arange(0, n_cycles * 20 + 1, 20)],
                                random_location_density=0.3, error_sd=1.0, random_seed=5,
                                stationary_observers=True, store_as_jax=True).observe()
    model = dab.model.Lorenz96Model(model_obj=dab.data.Lorenz96(system_dim=N, delta_t=0.01, store_as_jax=True))
    init = nature.isel(time=0)
    rng = np.random.default_rng(3)
    init = init.assign(x=(['ensemble', 'index'], init['x'].values + rng.standard_normal((k, N))))
    kw = dict(input_state=init, start_time=0.0, obs_vector=obs, n_cycles=n_cycles, obs_error_sd=1.0,
              analysis_window=0.2, analysis_time_in_window=0.0)
    sharded = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model, distributed=True).cycle(**kw)
    ok = True
    if rank == 0:
        single = dab.dacycler.ETKF(system_dim=N, delta_t=0.01, ensemble_dim=k, model_obj=model, distributed=False).cycle(**kw)
        a, b = sharded['x'].values, single['x'].values
        err = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        ok = a.shape == b.shape == (n_cycles, k, N) and bool(np.isfinite(a).all()) and err < 1e-10
        print("multigpu_api_check world=%d shape=%s rel_err=%.3e %s" % (world, a.shape, err, "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
