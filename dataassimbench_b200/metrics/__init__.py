"""Deterministic scores (mirrors dabench/metrics/_deterministic.py:20-89); host-side NumPy,
used to evaluate the 1 % time-mean-RMSE parity criterion."""
import numpy as np

__all__ = ['pearson_r', 'rmse', 'mse', 'mae', 'rmse_from_sse']


def mse(predictions, targets):
    d = np.asarray(predictions) - np.asarray(targets)
    return np.mean(d * d)


def rmse(predictions, targets):
    return np.sqrt(mse(predictions, targets))


def mae(predictions, targets):
    return np.mean(np.abs(np.asarray(predictions) - np.asarray(targets)))


def pearson_r(predictions, targets):
    p = np.asarray(predictions) - np.mean(predictions)
    t = np.asarray(targets) - np.mean(targets)
    return np.sum(t * p) / (np.sqrt(np.sum(t * t)) * np.sqrt(np.sum(p * p)))


def rmse_from_sse(sse, n):
    """RMSE from the per-cycle sums of squared errors the cycler accumulates on the device
    (`dab_cycler_set_truth`; summed over the ranks first when the grid is sharded)."""
    return np.sqrt(np.asarray(sse, dtype=np.float64) / n)
