"""Oracle: analysis-window bookkeeping (TEST INFRASTRUCTURE ONLY).

Integer/host part of `DACycler.cycle`; restates
  * _get_all_times     -- dabench/dacycler/_utils.py:14-38
  * _get_obs_indices   -- dabench/dacycler/_utils.py:41-84
  * _pad_time_indices  -- dabench/dacycler/_utils.py:87-119
exactly as written there (including the `isclose(rtol=0)` boundary handling, which
with the numpy/jax default atol = 1e-8 makes the window [t_c - w/2, t_c + w/2)
for the ETKF: start inclusive, end exclusive, dabench/dacycler/_dacycler.py:253-254).
"""
import numpy as np


def get_all_times(start_time, analysis_window, analysis_cycles):
    return (np.repeat(start_time, analysis_cycles)
            + np.arange(0, analysis_cycles * analysis_window, analysis_window))


def get_obs_indices(analysis_times, obs_times, analysis_window,
                    start_inclusive=True, end_inclusive=False):
    obs_times = np.asarray(obs_times, dtype=np.float64)
    out = []
    for cur_time in analysis_times:
        lo = cur_time - analysis_window / 2
        hi = cur_time + analysis_window / 2
        at_lo = np.isclose(obs_times, lo, rtol=0)
        at_hi = np.isclose(obs_times, hi, rtol=0)
        sel = ((obs_times > lo) * (obs_times < hi)
               * (1 - (1 - start_inclusive) * at_lo)
               * (1 - (1 - end_inclusive) * at_hi)
               + start_inclusive * at_lo
               + end_inclusive * at_hi)
        out.append(np.where(sel)[0])
    return out


def pad_time_indices(obs_indices, add_one=True):
    """[n_cycles, T_pad] int; 0 == "no observation time in this slot" (add_one=True)."""
    row_length = max(len(r) for r in obs_indices)
    padded = np.zeros((len(obs_indices), row_length), dtype=np.int64)
    for c, row in enumerate(obs_indices):
        padded[c, :len(row)] = np.asarray(row, dtype=np.int64) + int(add_one)
    return padded
