"""Hyper-parameter optimisation, host side (commons/GaussianProcessCommons.scala:66-92 `optimizeHypers`).

Expert grouping as the reference (GPC:26-31: E = Math.round(N / n_e), point i -> expert i % E); the BCM objective
(sum over experts of the negative log marginal likelihood, GPR:55-68, + gradient) runs on the GPU (`sgp_bcm_nll`);
the bound-constrained quasi-Newton loop is SciPy's L-BFGS-B in place of Breeze's `LBFGSB(lower, upper, maxIter,
tolerance = tol)` -- the reference pins no test on the iterate trajectory, so results are compared on the objective
value and on predictive metrics, not on theta."""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

from .engine import ProjectedProcessEngine


def group_for_experts(n_points: int, dataset_size_for_expert: int):
    """GPC:26-31."""
    n_experts = int(np.floor(n_points / dataset_size_for_expert + 0.5))
    if n_experts <= 0:
        raise ZeroDivisionError("numberOfExperts == 0 (N < datasetSizeForExpert / 2)")
    return [np.arange(e, n_points, n_experts) for e in range(n_experts)]


def pack_experts(X, y, dataset_size_for_expert: int):
    """Expert-major copy of (X, y) for sgp_experts_upload: expert e = points e, e+E, e+2E, ... (GPC:26-31).
    The grouping is a strided layout, so the bulk of the copy is one reshape/transpose (k = N // E full rounds);
    the first N % E experts get one more point each."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, d = X.shape
    n_experts = int(np.floor(n / dataset_size_for_expert + 0.5))           # Math.round(N / n_e), GPC:27
    if n_experts <= 0:
        raise ZeroDivisionError("numberOfExperts == 0 (N < datasetSizeForExpert / 2)")
    k, r = divmod(n, n_experts)                       # every expert has k points, the first r experts k + 1
    sizes = np.full(n_experts, k, dtype=np.int64)
    sizes[:r] += 1
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    Xp = np.empty((n, d)); yp = np.empty(n)
    if r == 0:
        Xp.reshape(n_experts, k, d)[...] = X.reshape(k, n_experts, d).transpose(1, 0, 2)
        yp.reshape(n_experts, k)[...] = y.reshape(k, n_experts).T
    else:
        # experts < r: k+1 points (rows e, e+E, ..., e+kE); experts >= r: k points
        head_X = X[:k * n_experts].reshape(k, n_experts, d)
        head_y = y[:k * n_experts].reshape(k, n_experts)
        a = Xp[:offsets[r]].reshape(r, k + 1, d)
        a[:, :k] = head_X[:, :r].transpose(1, 0, 2)
        a[:, k] = X[k * n_experts:]
        b = yp[:offsets[r]].reshape(r, k + 1)
        b[:, :k] = head_y[:, :r].T
        b[:, k] = y[k * n_experts:]
        Xp[offsets[r]:].reshape(n_experts - r, k, d)[...] = head_X[:, r:].transpose(1, 0, 2)
        yp[offsets[r]:].reshape(n_experts - r, k)[...] = head_y[:, r:].T
    return Xp, yp, offsets


class BcmObjective:
    """f(theta) and its gradient, memoised like commons/util/DiffFunctionMemoized.scala."""

    def __init__(self, engine: ProjectedProcessEngine, kernel_factory, X, y, dataset_size_for_expert: int):
        self.engine, self.kernel_factory = engine, kernel_factory
        # expert grouping (GPC:26-31) as a strided gather on the device while the points stream in (the host-side
        # pack_experts below is the same permutation, kept for callers that hold expert-major data and for the tests)
        self.n_experts = engine.experts_upload_grouped(X, y, dataset_size_for_expert)
        self._memo = {}
        self.evaluations = 0

    def __call__(self, theta):
        key = tuple(np.asarray(theta, dtype=np.float64))
        if key not in self._memo:
            kernel = self.kernel_factory().setHyperparameters(np.asarray(theta, dtype=np.float64))
            self._memo[key] = self.engine.bcm_nll(kernel)
            self.evaluations += 1
        return self._memo[key]


def optimize_hypers(gp, X, y, engine: ProjectedProcessEngine | None = None):
    """Returns the optimal hyper-parameter vector (layout of Kernel.getHyperparameters)."""
    own = engine is None
    engine = engine or ProjectedProcessEngine.acquire(gp._device)
    try:
        obj = BcmObjective(engine, gp.getKernel, X, y, gp._datasetSizeForExpert)
        k0 = gp.getKernel()
        x0 = k0.getHyperparameters()
        lo, up = k0.hyperparameterBoundaries()
        bounds = [(float(l), None if np.isinf(u) else float(u)) for l, u in zip(lo, up)]
        res = minimize(lambda t: obj(t), x0, jac=True, method="L-BFGS-B", bounds=bounds,
                       options=dict(maxiter=gp._maxIter, ftol=gp._tol, gtol=gp._tol))
        gp.last_objective = dict(value=float(res.fun), evaluations=obj.evaluations, iterations=int(res.nit),
                                 n_experts=obj.n_experts)
        return np.asarray(res.x, dtype=np.float64)
    finally:
        if own:
            engine.release()
