"""Oracle: deterministic scores (TEST INFRASTRUCTURE ONLY).
dabench/metrics/_deterministic.py:20-89."""
import numpy as np


def mse(predictions, targets):
    return np.mean(np.square(np.asarray(predictions) - np.asarray(targets)))


def rmse(predictions, targets):
    return np.sqrt(mse(predictions, targets))


def mae(predictions, targets):
    return np.mean(np.abs(np.asarray(predictions) - np.asarray(targets)))


def pearson_r(predictions, targets):
    t = np.asarray(targets) - np.mean(targets)
    p = np.asarray(predictions) - np.mean(predictions)
    return np.sum(t * p) / (np.sqrt(np.sum(t * t)) * np.sqrt(np.sum(p * p)))
