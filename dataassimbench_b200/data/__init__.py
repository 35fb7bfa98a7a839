from ._lorenz96 import Lorenz96, Data
from ._lorenz63 import Lorenz63

__all__ = ["Lorenz96", "Lorenz63", "Data"]
