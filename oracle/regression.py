"""Per-expert regression NLL + gradient (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows `regression/GaussianProcessRegression.scala:55-68` (GPR) and
`commons/util/logDetAndInv.scala:36-63`.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


class MatrixSingularException(Exception):
    pass


def log_det_and_inv(X: np.ndarray):
    """logDetAndInv.scala:58-63: one LU (dgetrf), then (sign, log|det|, inverse via dgetri)."""
    lu, piv = sla.lu_factor(X)                       # LU.primitive
    # LU2logdet :36-51 -- piv is 0-based here; counts rows actually exchanged.
    num_exchanged = int(np.sum(piv != np.arange(len(piv))))
    sign = -1.0 if num_exchanged % 2 == 1 else 1.0
    d = np.diag(lu)
    if np.any(d == 0.0):
        return 0.0, -np.inf, None
    logdet = float(np.sum(np.log(np.abs(d))))
    sign *= float(np.prod(np.sign(d)))
    getri, = sla.get_lapack_funcs(("getri",), (lu,))
    inv, info = getri(lu, piv)                       # LU2inv :14-30
    if info > 0:
        raise MatrixSingularException()
    return sign, logdet, inv


def regression_likelihood_and_gradient(y: np.ndarray, kernel, x: np.ndarray):
    """GPR:55-68.  nll = 1/2 y^T K^-1 y + 1/2 log|det K|  (no n/2 log 2pi term, sign of det ignored);
    grad_i = -1/2 sum(dK_i o (alpha alpha^T - K^-1))."""
    kernel.set_hyperparameters(x)
    k, derivative = kernel.training_kernel_and_derivative()
    _, logdet, kinv = log_det_and_inv(k)
    alpha = kinv @ y
    likelihood = 0.5 * float(y @ alpha) + 0.5 * logdet
    aat_minus_kinv = np.outer(alpha, alpha) - kinv
    gradient = np.array([-0.5 * float(np.sum(d * aat_minus_kinv)) for d in derivative])
    return likelihood, gradient


def bcm_objective(experts, x: np.ndarray):
    """GPC:73-78: sum of the per-expert (likelihood, gradient) -- the treeAggregate."""
    total, grad = 0.0, np.zeros(len(x))
    for y, k in experts:
        l, g = regression_likelihood_and_gradient(y, k, x)
        total += l
        grad += g
    return total, grad
