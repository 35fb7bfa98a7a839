from ._model import Model, Lorenz96Model

__all__ = ["Model", "Lorenz96Model"]
