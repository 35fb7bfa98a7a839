from ._dacycler import DACycler
from ._etkf import ETKF
from ._var3d import Var3D
from ._batched import cycle_batched

__all__ = ["DACycler", "ETKF", "Var3D", "cycle_batched"]
