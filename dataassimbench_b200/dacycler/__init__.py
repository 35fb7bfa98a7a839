from ._dacycler import DACycler
from ._etkf import ETKF

__all__ = ["DACycler", "ETKF"]
