"""CPU oracle for the DataAssimBench ETKF cycling hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/SciPy FP64 restatement of
the reference algorithm (StevePny/DataAssimBench, `dabench.dacycler.ETKF.cycle`
on a Lorenz-96 ensemble).  It exists so that the CUDA path can be checked for
parity; it is never the thing that is shipped or measured.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  Nothing under `dataassimbench_b200/` imports it.

Parity status: PINNED.  The restatement reproduces every known-answer vector
the reference's own tests hold for this path (see `tests/test_oracle_golden.py`):

* `tests/dacycler_etkf_test.py:89-102`  -- three end-to-end ETKF 5-vectors
* `tests/data_lorenz96_test.py:120-136` -- Lorenz-96 trajectory
* `tests/observer_base_test.py:133-180` -- observer indices / times / errors
* `tests/metrics_deterministic.py`      -- rmse / mse / mae

The reference itself (JAX + xarray + xarray_jax) cannot be imported in this
image, so there is no `oracle/_ref`; third-party arithmetic the reference
delegates to (`jax.experimental.ode.odeint`, `jax.random.normal`,
`jnp.linalg.pinv`, `jax.scipy.linalg.sqrtm`) is restated from the published
algorithms in `odeint_dopri5.py`, `threefry.py` and `etkf.py`.
"""

from . import l96, odeint_dopri5, windows, etkf, observer, threefry, metrics  # noqa: F401
