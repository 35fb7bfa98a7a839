from ._model import Model, Lorenz96Model, Lorenz63Model, SQGTurbModel

__all__ = ["Model", "Lorenz96Model", "Lorenz63Model", "SQGTurbModel"]
