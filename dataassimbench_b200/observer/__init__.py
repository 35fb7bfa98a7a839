from ._observer import Observer

__all__ = ["Observer"]
