"""Process-wide native handle (one per GPU) for the host-side API objects."""
from . import _native

_handles = {}


def handle(device=None):
    """The `dab_handle` of `device` (default: torch's current CUDA device).  Raises if the CUDA
    library or a B200 is missing: there is no CPU fallback."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("dataassimbench_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    h = _handles.get(device)
    if h is None:
        h = _native.Handle(device)
        _handles[device] = h
    return h
