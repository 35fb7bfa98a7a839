from ._lorenz96 import Lorenz96, Data
from ._lorenz63 import Lorenz63
from ._sqgturb import SQGTurb

__all__ = ["Lorenz96", "Lorenz63", "SQGTurb", "Data"]
