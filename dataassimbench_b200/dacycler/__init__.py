from ._dacycler import DACycler
from ._etkf import ETKF
from ._batched import cycle_batched

__all__ = ["DACycler", "ETKF", "cycle_batched"]
