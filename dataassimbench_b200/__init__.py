"""dataassimbench_b200 -- B200-native ETKF cycling hot path for DataAssimBench.

Same Python surface as the reference for this one path (`dacycler.ETKF.cycle()`, `model.Model`,
`data.Lorenz96`, `observer.Observer`, `obsop`, `vector`); the arithmetic runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/dab_b200.h.  No CPU fallback.
"""
__version__ = "0.1.0"

from . import data, model, dacycler, observer, obsop, vector, metrics  # noqa: E402,F401
