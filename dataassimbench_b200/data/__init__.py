from ._lorenz96 import Lorenz96, Data

__all__ = ["Lorenz96", "Data"]
