"""Host-side mirror of the reference's Estimator surface for regression
(`regression/GaussianProcessRegression.scala`, `commons/GaussianProcessParams.scala`,
`commons/GaussianProcessCommons.scala`): the same setters, defaults and call order; the data-parallel
statistics, the m x m tail and prediction run in the CUDA library.

Spark `Dataset` plumbing (label/features columns, RDD caching) has no meaning here: `fit` takes the
feature matrix and the labels directly."""
from __future__ import annotations

import numpy as np

from . import _native as N
from .engine import ProjectedProcessEngine, OperandRangeError
from .kernels import Kernel, RBFKernel, EyeKernel, const


class ActiveSetProvider:
    """commons/ActiveSetProvider.scala:13-20."""
    def __call__(self, activeSetSize, X, y, kernel_factory, optimalHyperparameter, seed):
        raise NotImplementedError


class _RandomActiveSetProvider(ActiveSetProvider):
    """commons/ActiveSetProvider.scala:48-56: uniform sample without replacement.  (Spark's `takeSample`
    RNG stream is not reproduced -- the reference pins no test on it; pass an explicit active set through
    `ExplicitActiveSetProvider` when results must be compared.)"""
    def __call__(self, activeSetSize, X, y, kernel_factory, optimalHyperparameter, seed):
        rng = np.random.default_rng(seed)
        idx = rng.choice(len(X), size=min(activeSetSize, len(X)), replace=False)
        return np.asarray(X, dtype=np.float64)[idx]


RandomActiveSetProvider = _RandomActiveSetProvider()


class ExplicitActiveSetProvider(ActiveSetProvider):
    def __init__(self, active_set):
        self.active_set = np.asarray(active_set, dtype=np.float64)

    def __call__(self, activeSetSize, X, y, kernel_factory, optimalHyperparameter, seed):
        return self.active_set


class GaussianProcessParams:
    """commons/GaussianProcessParams.scala:8-54 (defaults at :33-51)."""

    def __init__(self):
        self._activeSetProvider = RandomActiveSetProvider
        self._kernel = lambda: RBFKernel()
        self._datasetSizeForExpert = 100
        self._sigma2 = 1e-3
        self._activeSetSize = 100
        self._maxIter = 100
        self._tol = 1e-6
        self._seed = 0

    def setActiveSetProvider(self, value): self._activeSetProvider = value; return self
    def setDatasetSizeForExpert(self, value: int): self._datasetSizeForExpert = int(value); return self
    def setMaxIter(self, value: int): self._maxIter = int(value); return self
    def setSigma2(self, value: float): self._sigma2 = float(value); return self
    def setKernel(self, value): self._kernel = value; return self
    def setTol(self, value: float): self._tol = float(value); return self
    def setActiveSetSize(self, value: int): self._activeSetSize = int(value); return self
    def setSeed(self, value: int): self._seed = int(value); return self

    def getKernel(self) -> Kernel:
        """commons/GaussianProcessCommons.scala:18: user kernel + sigma2.const * EyeKernel."""
        return self._kernel() + const(self._sigma2) * EyeKernel()


class GaussianProjectedProcessRawPredictor:
    """commons/GaussianProcessCommons.scala:118-126.  Holds the engine whose device memory carries the
    active set, magicVector and magicMatrix."""

    def __init__(self, engine: ProjectedProcessEngine, magicVector, magicMatrix, kernel: Kernel, activeSet):
        self._engine = engine
        self.magicVector, self.magicMatrix, self.kernel, self.activeSet = magicVector, magicMatrix, kernel, activeSet

    def predict(self, features):
        """(mean, variance) for one vector or a block of vectors."""
        f = np.asarray(features, dtype=np.float64)
        mean, var = self._engine.predict(f, with_variance=True)
        if f.ndim == 1:
            return float(mean[0]), float(var[0])
        return mean, var


class GaussianProcessRegressionModel:
    """regression/GaussianProcessRegression.scala:75-87: `predict` returns the mean only."""

    def __init__(self, rawPredictor: GaussianProjectedProcessRawPredictor, hyperparameters):
        self.rawPredictor = rawPredictor
        self.hyperparameters = hyperparameters

    def predict(self, features):
        f = np.asarray(features, dtype=np.float64)
        mean, _ = self.rawPredictor._engine.predict(f, with_variance=False)
        return float(mean[0]) if f.ndim == 1 else mean

    transform = predict


class GaussianProcessRegression(GaussianProcessParams):
    """regression/GaussianProcessRegression.scala:36-73."""

    def __init__(self, device: int = 0, shard_points: int = 1 << 22):
        super().__init__()
        self._device = device
        self._shard_points = shard_points
        self.last_stats = None

    def fit(self, X, y, hyperparameters=None) -> GaussianProcessRegressionModel:
        """GPR.train (:43-53).  `hyperparameters` short-circuits optimizeHypers (:48) with a given theta;
        otherwise the BCM objective is optimised when maxIter > 0."""
        X = np.asarray(X)
        y = np.asarray(y, dtype=np.float64)
        if hyperparameters is None:
            if self._maxIter > 0:
                from .hyperopt import optimize_hypers     # per-expert BCM NLL on the GPU + L-BFGS-B on the host
                hyperparameters = optimize_hypers(self, X, y)
            else:
                hyperparameters = self.getKernel().getHyperparameters()
        theta = np.asarray(hyperparameters, dtype=np.float64)
        return self._produce_model(X, y, theta)

    def _produce_model(self, X, y, theta):
        """produceModel -> projectedProcess  (commons/GaussianProcessCommons.scala:40-59, 102-110)."""
        active_set = self._activeSetProvider(self._activeSetSize, X, y, self.getKernel, theta, self._seed)
        kernel = self.getKernel().setHyperparameters(theta)
        eng = ProjectedProcessEngine(self._device)

        def stats():
            eng.begin(kernel, active_set)                               # PGPH:23 broadcast(activeSet)
            for s in range(0, len(X), self._shard_points):              # PGPH:25-35 seqOp over shards
                eng.accumulate(X[s:s + self._shard_points], y[s:s + self._shard_points])
            return eng.finish()
        try:
            G, b = stats()
        except OperandRangeError:                                       # still on the GPU: fp64 DMMA kernel
            eng.set_precision(N.SGP_PREC_F64)
            G, b = stats()
        mv, mm = eng.magic()                                            # PGPH:49-60
        self.last_stats = (G, b)
        return GaussianProcessRegressionModel(GaussianProjectedProcessRawPredictor(eng, mv, mm, kernel, active_set),
                                              theta)
