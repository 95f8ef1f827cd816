"""Small host-side helpers of the reference's examples."""
from __future__ import annotations

import numpy as np


def scale(X) -> np.ndarray:
    """commons/util/Scaling.scala:10-25 (used by regression/examples/Airfoil.scala:26-32 and
    classification/examples/MNIST.scala:42-45): subtract the column mean, divide by the POPULATION standard deviation
    (variance / n); a zero variance counts as 1."""
    X = np.asarray(X, dtype=np.float64)
    n = float(len(X))
    mean = X.sum(axis=0) / n
    centered = X - mean
    variance = (centered * centered).sum(axis=0) / n
    return centered / np.sqrt(np.where(variance > 0.0, variance, 1.0))
