"""Host-side mirror of the reference's kernel DSL (`commons/kernel/*.scala`): same class names, the same
hyperparameter layout and bounds, the same sugar -- but no numerics.  A kernel object here is a
*description*; `flatten()` turns the tree into the term list of `sgp_kernel_desc` (include/sgp.h) and
all arithmetic happens in the CUDA library.

Scala                                   Python
  1 * new ARDRBFKernel(5)                 1 * ARDRBFKernel(5)
  1.const * new EyeKernel                 const(1) * EyeKernel()
  (0.5 between 0 and 1) * new EyeKernel   Scalar(0.5).between(0).and_(1) * EyeKernel()
  k1 + k2                                 k1 + k2
"""
from __future__ import annotations

import numpy as np

from . import _native as N


class Kernel:
    """commons/kernel/Kernel.scala:12-98 -- the parts that are pure bookkeeping."""

    def getHyperparameters(self) -> np.ndarray: raise NotImplementedError
    def setHyperparameters(self, value): raise NotImplementedError
    def numberOfHyperparameters(self) -> int: raise NotImplementedError
    def hyperparameterBoundaries(self): raise NotImplementedError
    def flatten(self, scale: float = 1.0) -> list: raise NotImplementedError

    def hyper_descriptors(self, scale: float = 1.0, term_offset: int = 0) -> list:
        """One dict per hyper-parameter, in getHyperparameters order, saying how the kernel matrix depends on it
        (consumed by sgp_bcm_nll): {"kind": SCALE, "coef": {term_index: d scale_t / d theta}} for a trainable scalar,
        {"kind": ARD_BETA, "term": t, "dim": k, "value": beta_k}, {"kind": RBF_SIGMA, "term": t, "value": sigma}."""
        raise NotImplementedError

    @property
    def whiteNoiseVar(self) -> float:
        return sum(t["scale"] for t in self.flatten() if t["type"] == N.SGP_TERM_EYE)

    def __add__(self, other): return SumOfKernels(self, other)

    def __rmul__(self, c):
        return (c if isinstance(c, Scalar) else Scalar(float(c))) * self


class EyeKernel(Kernel):
    """Kernel.scala:142-164."""
    def getHyperparameters(self): return np.zeros(0)
    def setHyperparameters(self, value): return self
    def numberOfHyperparameters(self): return 0
    def hyperparameterBoundaries(self): return np.zeros(0), np.zeros(0)
    def flatten(self, scale=1.0): return [dict(type=N.SGP_TERM_EYE, scale=scale)]
    def hyper_descriptors(self, scale=1.0, term_offset=0): return []
    def __str__(self): return "I"


class ARDRBFKernel(Kernel):
    """ARDRBFKernel.scala:20-96; `ARDRBFKernel(p)` with an int is the `this(p: Int, beta = 1, ...)` ctor."""

    def __init__(self, beta, lower=None, upper=None):
        if isinstance(beta, (int, np.integer)):
            p = int(beta)
            self.beta = np.ones(p)
            self.lower = np.zeros(p) if lower is None else np.zeros(p) + lower
            self.upper = np.full(p, np.inf) if upper is None else np.zeros(p) + upper
        else:
            self.beta = np.array(beta, dtype=np.float64)
            self.lower = self.beta * 0.0 if lower is None else np.array(lower, dtype=np.float64)
            self.upper = np.full(len(self.beta), np.inf) if upper is None else np.array(upper, dtype=np.float64)

    def getHyperparameters(self): return self.beta
    def setHyperparameters(self, value):
        self.beta = np.array(value, dtype=np.float64)
        return self
    def numberOfHyperparameters(self): return len(self.beta)
    def hyperparameterBoundaries(self): return self.lower, self.upper
    def flatten(self, scale=1.0): return [dict(type=N.SGP_TERM_ARD, scale=scale, beta=self.beta.copy())]
    def hyper_descriptors(self, scale=1.0, term_offset=0):
        return [dict(kind=N.SGP_HYPER_ARD_BETA, term=term_offset, dim=k, value=float(b)) for k, b in enumerate(self.beta)]
    def __str__(self): return "ARDRBFKernel(beta=[" + ", ".join("%1.1e" % e for e in self.beta) + "])"


class RBFKernel(Kernel):
    """RBFKernel.scala:14-85."""

    def __init__(self, sigma: float = 1.0, lower: float = 1e-6, upper: float = np.inf):
        self.sigma, self.lower, self.upper = float(sigma), float(lower), float(upper)

    def getHyperparameters(self): return np.array([self.sigma])
    def setHyperparameters(self, value):
        self.sigma = float(np.asarray(value, dtype=np.float64)[0])
        return self
    def numberOfHyperparameters(self): return 1
    def hyperparameterBoundaries(self): return np.array([self.lower]), np.array([self.upper])
    def flatten(self, scale=1.0): return [dict(type=N.SGP_TERM_RBF, scale=scale, sigma=self.sigma)]
    def hyper_descriptors(self, scale=1.0, term_offset=0):
        return [dict(kind=N.SGP_HYPER_RBF_SIGMA, term=term_offset, dim=0, value=self.sigma)]
    def __str__(self): return "RBFKernel(sigma=%1.1e)" % self.sigma


class ConstantTimesKernel(Kernel):
    """ScalarTimesKernel.scala:41-59."""

    def __init__(self, kernel: Kernel, C: float):
        if not C >= 0:
            raise ValueError("requirement failed: C should be positive")
        self.kernel, self.C = kernel, float(C)

    def getHyperparameters(self): return self.kernel.getHyperparameters()
    def setHyperparameters(self, value):
        self.kernel.setHyperparameters(value)
        return self
    def numberOfHyperparameters(self): return self.kernel.numberOfHyperparameters()
    def hyperparameterBoundaries(self): return self.kernel.hyperparameterBoundaries()
    def flatten(self, scale=1.0): return self.kernel.flatten(scale * self.C)
    def hyper_descriptors(self, scale=1.0, term_offset=0):
        return self.kernel.hyper_descriptors(scale * self.C, term_offset)
    def __str__(self): return ("%1.1e * %s" % (self.C, self.kernel)) if self.C != 0 else ""


class TrainableScalarTimesKernel(ConstantTimesKernel):
    """ScalarTimesKernel.scala:71-98: C is a hyperparameter, PREPENDED to the inner kernel's vector."""

    def __init__(self, kernel: Kernel, C: float, Clower: float = 0.0, Cupper: float = np.inf):
        super().__init__(kernel, C)
        self.Clower, self.Cupper = float(Clower), float(Cupper)

    def getHyperparameters(self): return np.concatenate([[self.C], self.kernel.getHyperparameters()])
    def setHyperparameters(self, value):
        value = np.asarray(value, dtype=np.float64)
        self.C = float(value[0])
        self.kernel.setHyperparameters(value[1:])
        return self
    def numberOfHyperparameters(self): return 1 + self.kernel.numberOfHyperparameters()
    def hyperparameterBoundaries(self):
        lo, up = self.kernel.hyperparameterBoundaries()
        return np.concatenate([[self.Clower], lo]), np.concatenate([[self.Cupper], up])
    def hyper_descriptors(self, scale=1.0, term_offset=0):
        # dK/dC = (outer scale) * inner kernel matrix  (ScalarTimesKernel.scala:93-97): per inner term, its scale without C
        inner = self.kernel.flatten(scale)
        own = dict(kind=N.SGP_HYPER_SCALE, coef={term_offset + i: t["scale"] for i, t in enumerate(inner)})
        return [own] + self.kernel.hyper_descriptors(scale * self.C, term_offset)


class Scalar:
    """ScalarTimesKernel.scala:100-141."""

    def __init__(self, C, lower=0.0, upper=np.inf, isTrainable=True):
        if not ((lower < upper and isTrainable) or not isTrainable):
            raise ValueError("The scalar should either have its lower limit below its upper limit "
                             "or not be trainable")
        self.C, self.lower, self.upper, self.isTrainable = float(C), float(lower), float(upper), isTrainable

    def __mul__(self, kernel: Kernel):
        if self.isTrainable:
            return TrainableScalarTimesKernel(kernel, self.C, self.lower, self.upper)
        return ConstantTimesKernel(kernel, self.C)

    def between(self, lower):
        outer = self

        class _And:
            def and_(self, upper): return Scalar(outer.C, lower, upper, outer.isTrainable)
        return _And()

    def below(self, newUpper): return Scalar(self.C, self.lower, newUpper, self.isTrainable)

    @property
    def const(self): return Scalar(self.C, self.C, self.C, False)


def const(c: float) -> Scalar:
    return Scalar(c).const


def WhiteNoiseKernel(initial, lower, upper) -> Kernel:
    """Kernel.scala:166-169."""
    return Scalar(initial).between(lower).and_(upper) * EyeKernel()


class SumOfKernels(Kernel):
    """SumOfKernels.scala:15-65: hyperparameters are concatenated left to right."""

    def __init__(self, kernel1: Kernel, kernel2: Kernel):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def getHyperparameters(self):
        return np.concatenate([self.kernel1.getHyperparameters(), self.kernel2.getHyperparameters()])
    def setHyperparameters(self, value):
        value = np.asarray(value, dtype=np.float64)
        n1 = self.kernel1.numberOfHyperparameters()
        self.kernel1.setHyperparameters(value[:n1])
        self.kernel2.setHyperparameters(value[n1:])
        return self
    def numberOfHyperparameters(self):
        return self.kernel1.numberOfHyperparameters() + self.kernel2.numberOfHyperparameters()
    def hyperparameterBoundaries(self):
        l1, u1 = self.kernel1.hyperparameterBoundaries()
        l2, u2 = self.kernel2.hyperparameterBoundaries()
        return np.concatenate([l1, l2]), np.concatenate([u1, u2])
    def flatten(self, scale=1.0): return self.kernel1.flatten(scale) + self.kernel2.flatten(scale)
    def hyper_descriptors(self, scale=1.0, term_offset=0):
        n1 = len(self.kernel1.flatten(scale))
        return (self.kernel1.hyper_descriptors(scale, term_offset)
                + self.kernel2.hyper_descriptors(scale, term_offset + n1))
    def __str__(self): return " + ".join(s for s in (str(self.kernel1), str(self.kernel2)) if len(s) > 0)
