"""Analysis-window bookkeeping (host, integer): which observation times feed which cycle.

Same results as the reference's `_get_all_times`, `_get_obs_indices` and `_pad_time_indices`
(dabench/dacycler/_utils.py:14-38, 41-84, 87-119), computed as one broadcast over
[cycles, observation times] instead of a Python list comprehension per cycle.
"""
import numpy as np

_ATOL = 1e-8      # numpy/jax `isclose(..., rtol=0)` default atol, used by the reference at the window edges


def window_centres(start_time, analysis_window, n_cycles):
    """`start + arange(0, n*w, w)`.  The reference builds the offsets with a float `arange`
    (_utils.py:32-36); for some (n, w) that array has n+1 elements and the reference fails on the
    broadcast -- the same ValueError is raised here rather than silently changing the grid."""
    offsets = np.arange(0, n_cycles * analysis_window, analysis_window)
    if len(offsets) != n_cycles:
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,): "
                         "arange(0, n_cycles*analysis_window, analysis_window) is not n_cycles long "
                         "(reference behaviour, dabench/dacycler/_utils.py:32-36)" % (n_cycles, len(offsets)))
    return start_time + offsets


def padded_time_table(obs_times, centres, analysis_window, start_inclusive=True, end_inclusive=False):
    """[n_cycles, T_pad] int64, 0 = empty slot, otherwise (observation-time index + 1): the
    `all_filtered_padded` array of dabench/dacycler/_dacycler.py:250-259."""
    t = np.asarray(obs_times, dtype=np.float64)[None, :]
    lo = (np.asarray(centres, dtype=np.float64) - analysis_window / 2)[:, None]
    hi = (np.asarray(centres, dtype=np.float64) + analysis_window / 2)[:, None]
    at_lo = np.abs(t - lo) <= _ATOL
    at_hi = np.abs(t - hi) <= _ATOL
    sel = (t > lo) & (t < hi)
    if not start_inclusive:
        sel &= ~at_lo
    if not end_inclusive:
        sel &= ~at_hi
    if start_inclusive:
        sel |= at_lo
    if end_inclusive:
        sel |= at_hi
    counts = sel.sum(axis=1)
    t_pad = int(counts.max()) if sel.size else 0
    table = np.zeros((sel.shape[0], t_pad), dtype=np.int64)
    rows, cols = np.nonzero(sel)                      # row-major: per cycle, increasing time index
    slot = np.arange(len(rows)) - np.repeat(np.cumsum(counts) - counts, counts)
    table[rows, slot] = cols + 1
    return table
