"""Feature standardisation (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows `commons/util/Scaling.scala:10-25`: subtract the mean, divide by the
POPULATION standard deviation (variance / n), zero variance -> 1.
"""
import numpy as np


def scale(X: np.ndarray) -> np.ndarray:
    X = np.asarray(X, dtype=np.float64)
    n = float(len(X))
    mean = X.sum(axis=0) / n
    centered = X - mean
    variance = (centered * centered).sum(axis=0) / n
    variance = np.where(variance > 0.0, variance, 1.0)
    return centered / np.sqrt(variance)
