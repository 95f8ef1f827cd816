"""Kernel DSL restatement (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows `commons/kernel/*.scala` of the reference.  Citations are relative to
`/root/reference/src/main/scala/org/apache/spark/ml/commons/kernel/`.
All arithmetic fp64, matrices are numpy arrays (orientation as in the
reference: `crossKernel(test)` is `len(test) x len(train)`, Kernel.scala:69-74).
"""
from __future__ import annotations

import math
import numpy as np


class TrainingVectorsNotInitializedException(Exception):
    """Kernel.scala:116-117."""

    def __init__(self):
        super().__init__("setTrainingVectors method should have been called first")


def _as2d(vectors) -> np.ndarray:
    a = np.asarray(vectors, dtype=np.float64)
    if a.ndim == 1:
        a = a[None, :]
    return a


class Kernel:
    """Kernel.scala:12-98 (trait Kernel)."""

    def get_hyperparameters(self) -> np.ndarray: raise NotImplementedError
    def set_hyperparameters(self, value): raise NotImplementedError
    def number_of_hyperparameters(self) -> int: raise NotImplementedError
    def hyperparameter_boundaries(self): raise NotImplementedError
    def get_training_vectors(self) -> np.ndarray: raise NotImplementedError
    def set_training_vectors(self, vectors): raise NotImplementedError
    def training_kernel(self) -> np.ndarray: raise NotImplementedError
    def training_kernel_diag(self) -> np.ndarray: raise NotImplementedError
    def training_kernel_and_derivative(self): raise NotImplementedError
    def cross_kernel(self, test) -> np.ndarray: raise NotImplementedError
    def self_kernel(self, test) -> float: raise NotImplementedError

    @property
    def white_noise_var(self) -> float: raise NotImplementedError

    def cross_kernel_vec(self, test) -> np.ndarray:
        """Kernel.scala:81-84: single-vector overload, returns the row k(test, train_j)."""
        return self.cross_kernel(_as2d(test))[0, :]

    # package.scala:6-8 (`+`) and package.scala:4 + ScalarTimesKernel.scala:108 (`Double * Kernel`)
    def __add__(self, other: "Kernel") -> "SumOfKernels":
        return SumOfKernels(self, other)

    def __rmul__(self, c) -> "Kernel":
        if isinstance(c, Scalar):
            return c * self
        return Scalar(float(c)) * self


class _TrainDatasetBearing(Kernel):
    """Kernel.scala:123-133."""

    def __init__(self):
        self._train = None

    def get_training_vectors(self):
        if self._train is None:
            raise TrainingVectorsNotInitializedException()
        return self._train

    def set_training_vectors(self, vectors):
        self._train = _as2d(vectors)
        return self


class EyeKernel(_TrainDatasetBearing):
    """Kernel.scala:142-164."""

    def get_hyperparameters(self): return np.zeros(0)
    def set_hyperparameters(self, value): return self
    def number_of_hyperparameters(self): return 0
    def hyperparameter_boundaries(self): return np.zeros(0), np.zeros(0)

    def training_kernel(self):                      # Kernel.scala:151
        return np.eye(len(self.get_training_vectors()))

    def training_kernel_diag(self):                 # Kernel.scala:111-114
        return np.zeros(len(self.get_training_vectors())) + 1.0

    def training_kernel_and_derivative(self):       # Kernel.scala:153-155
        return self.training_kernel(), []

    def cross_kernel(self, test):                   # Kernel.scala:157 -- zeros!
        return np.zeros((len(_as2d(test)), len(self.get_training_vectors())))

    @property
    def white_noise_var(self): return 1.0           # Kernel.scala:159

    def self_kernel(self, test): return 1.0         # Kernel.scala:161

    def __str__(self): return "I"


class ARDRBFKernel(_TrainDatasetBearing):
    """ARDRBFKernel.scala:20-96.  k(a,b) = exp(-||(a-b) o beta||^2)  (no 1/2; beta is an
    inverse length-scale)."""

    def __init__(self, beta, lower=None, upper=None):
        super().__init__()
        if np.isscalar(beta) and isinstance(beta, (int, np.integer)) and lower is None and upper is None:
            # this(p: Int, beta = 1, lower = 0, upper = inf)  ARDRBFKernel.scala:27-30
            p = int(beta)
            self.beta = np.zeros(p) + 1.0
            self.lower = np.zeros(p)
            self.upper = np.zeros(p) + np.inf
        else:
            self.beta = np.asarray(beta, dtype=np.float64).copy()
            # this(beta) = this(beta, beta*0, beta*inf)   ARDRBFKernel.scala:25
            self.lower = self.beta * 0.0 if lower is None else np.asarray(lower, dtype=np.float64)
            self.upper = self.beta * np.inf if upper is None else np.asarray(upper, dtype=np.float64)

    @classmethod
    def of_dim(cls, p: int, beta: float = 1.0, lower: float = 0.0, upper: float = np.inf):
        """ARDRBFKernel.scala:27-30."""
        return cls(np.zeros(p) + beta, np.zeros(p) + lower, np.zeros(p) + upper)

    def set_hyperparameters(self, value):
        self.beta = np.asarray(value, dtype=np.float64).copy()
        return self

    def get_hyperparameters(self): return self.beta
    def number_of_hyperparameters(self): return len(self.beta)
    def hyperparameter_boundaries(self): return self.lower, self.upper

    def _kernel_elements(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """ARDRBFKernel.scala:43-46 for every pair (a_i, b_j): norm((a-b)*:*beta), squared, exp(-.)."""
        # accumulated feature by feature (same terms, in index order, as breeze's norm loop); keeps the
        # temporaries at len(a) x len(b) instead of len(a) x len(b) x d
        acc = np.zeros((len(a), len(b)))
        for k in range(a.shape[1]):
            diff = (a[:, k, None] - b[None, :, k]) * self.beta[k]
            diff *= diff
            acc += diff
        wd = np.sqrt(acc)                            # breeze `norm` (2-norm) ...
        return np.exp(-wd * wd)                      # ... squared again (ARDRBFKernel.scala:44-45)

    def training_kernel(self):                      # ARDRBFKernel.scala:48-59
        t = self.get_training_vectors()
        return self._kernel_elements(t, t)

    def training_kernel_diag(self):
        return np.zeros(len(self.get_training_vectors())) + 1.0

    def training_kernel_and_derivative(self):       # ARDRBFKernel.scala:61-79
        t = self.get_training_vectors()
        K = self.training_kernel()
        minus2K = -2.0 * K
        diff = t[:, None, :] - t[None, :, :]
        beta_d2 = diff * diff * self.beta            # :68-70  (x_i-x_j)^2 * beta
        derivs = [beta_d2[:, :, k] * minus2K for k in range(len(self.beta))]
        return K, derivs

    def cross_kernel(self, test):                   # ARDRBFKernel.scala:81-89
        return self._kernel_elements(_as2d(test), self.get_training_vectors())

    @property
    def white_noise_var(self): return 0.0           # NoiselessKernel  Kernel.scala:103-105

    def self_kernel(self, test): return 1.0         # ARDRBFKernel.scala:91

    def __str__(self):
        return "ARDRBFKernel(beta=[" + ", ".join("%1.1e" % e for e in self.beta) + "])"


def _sqdist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """`Vectors.sqdist` for every pair: sum_k (a_ik - b_jk)^2 (direct form, as Spark does for dense)."""
    acc = np.zeros((len(a), len(b)))
    for k in range(a.shape[1]):
        diff = a[:, k, None] - b[None, :, k]
        diff *= diff
        acc += diff
    return acc


class RBFKernel(_TrainDatasetBearing):
    """RBFKernel.scala:14-85.  k = exp(-||x-z||^2 / (2 sigma^2))."""

    def __init__(self, sigma: float = 1.0, lower: float = 1e-6, upper: float = np.inf):
        super().__init__()
        self.sigma = float(sigma)
        self.lower = float(lower)
        self.upper = float(upper)
        self._sqd = None

    def set_hyperparameters(self, value):
        self.sigma = float(np.asarray(value, dtype=np.float64)[0])
        return self

    def get_hyperparameters(self): return np.array([self.sigma])
    def number_of_hyperparameters(self): return 1
    def hyperparameter_boundaries(self): return np.array([self.lower]), np.array([self.upper])

    def set_training_vectors(self, vectors):        # RBFKernel.scala:37-48
        super().set_training_vectors(vectors)
        self._sqd = _sqdist(self._train, self._train)
        return self

    def training_kernel(self):                      # RBFKernel.scala:50-54
        if self._sqd is None:
            raise TrainingVectorsNotInitializedException()
        return np.exp(self._sqd / (-2.0 * self.sigma * self.sigma))

    def training_kernel_diag(self):
        return np.zeros(len(self.get_training_vectors())) + 1.0

    def training_kernel_and_derivative(self):       # RBFKernel.scala:56-64
        if self._sqd is None:
            raise TrainingVectorsNotInitializedException()
        K = self.training_kernel()
        return K, [self._sqd * K / (self.sigma ** 3)]

    def cross_kernel(self, test):                   # RBFKernel.scala:66-76
        train = self.get_training_vectors()
        return np.exp(_sqdist(_as2d(test), train) / (-2.0 * self.sigma * self.sigma))

    @property
    def white_noise_var(self): return 0.0

    def self_kernel(self, test): return 1.0

    def __str__(self): return "RBFKernel(sigma=%1.1e)" % self.sigma


class _ScalarTimesKernel(Kernel):
    """ScalarTimesKernel.scala:6-31."""

    def __init__(self, kernel: Kernel, C: float):
        if not C >= 0:
            raise ValueError("requirement failed: C should be positive")   # :7
        self.kernel = kernel
        self.C = float(C)

    def get_training_vectors(self): return self.kernel.get_training_vectors()

    def set_training_vectors(self, vectors):
        self.kernel.set_training_vectors(vectors)
        return self

    def training_kernel(self): return self.kernel.training_kernel() * self.C          # :20
    def training_kernel_diag(self): return self.kernel.training_kernel_diag() * self.C  # :22
    def cross_kernel(self, test): return self.kernel.cross_kernel(test) * self.C      # :24
    def self_kernel(self, test): return self.kernel.self_kernel(test) * self.C        # :26

    @property
    def white_noise_var(self): return self.C * self.kernel.white_noise_var            # :28

    def __str__(self): return ("%1.1e * %s" % (self.C, self.kernel)) if self.C != 0 else ""


class ConstantTimesKernel(_ScalarTimesKernel):
    """ScalarTimesKernel.scala:41-59."""

    def get_hyperparameters(self): return self.kernel.get_hyperparameters()

    def set_hyperparameters(self, value):
        self.kernel.set_hyperparameters(value)
        return self

    def training_kernel_and_derivative(self):
        K, d = self.kernel.training_kernel_and_derivative()
        return K * self.C, [x * self.C for x in d]

    def number_of_hyperparameters(self): return self.kernel.number_of_hyperparameters()
    def hyperparameter_boundaries(self): return self.kernel.hyperparameter_boundaries()


class TrainableScalarTimesKernel(_ScalarTimesKernel):
    """ScalarTimesKernel.scala:71-98.  Hyperparameter vector = C prepended to the inner kernel's."""

    def __init__(self, kernel: Kernel, C: float, Clower: float = 0.0, Cupper: float = np.inf):
        super().__init__(kernel, C)
        self.Clower = float(Clower)
        self.Cupper = float(Cupper)

    def get_hyperparameters(self):
        return np.concatenate([[self.C], self.kernel.get_hyperparameters()])

    def set_hyperparameters(self, value):
        value = np.asarray(value, dtype=np.float64)
        self.C = float(value[0])
        self.kernel.set_hyperparameters(value[1:])
        return self

    def number_of_hyperparameters(self): return 1 + self.kernel.number_of_hyperparameters()

    def hyperparameter_boundaries(self):
        lo, up = self.kernel.hyperparameter_boundaries()
        return np.concatenate([[self.Clower], lo]), np.concatenate([[self.Cupper], up])

    def training_kernel_and_derivative(self):       # :93-97
        K, d = self.kernel.training_kernel_and_derivative()
        return K * self.C, [K] + [x * self.C for x in d]


class Scalar:
    """ScalarTimesKernel.scala:100-141 (the `1 between 0 and 30`, `1 below 10`, `1.const` sugar)."""

    def __init__(self, C: float, lower: float = 0.0, upper: float = np.inf, is_trainable: bool = True):
        if not ((lower < upper and is_trainable) or not is_trainable):
            raise ValueError("The scalar should either have its lower limit below its upper limit "
                             "or not be trainable")
        self.C, self.lower, self.upper, self.is_trainable = float(C), float(lower), float(upper), is_trainable

    def __mul__(self, kernel: Kernel) -> Kernel:
        if self.is_trainable:
            return TrainableScalarTimesKernel(kernel, self.C, self.lower, self.upper)
        return ConstantTimesKernel(kernel, self.C)

    def between(self, lower: float):
        outer = self

        class _And:
            def and_(self, upper: float):
                return Scalar(outer.C, lower, upper, outer.is_trainable)
        return _And()

    def below(self, new_upper: float): return Scalar(self.C, self.lower, new_upper, self.is_trainable)

    @property
    def const(self): return Scalar(self.C, self.C, self.C, False)


def const(c: float) -> Scalar:
    """`c.const`."""
    return Scalar(c).const


def WhiteNoiseKernel(initial: float, lower: float, upper: float) -> Kernel:
    """Kernel.scala:166-169."""
    return Scalar(initial).between(lower).and_(upper) * EyeKernel()


class SumOfKernels(Kernel):
    """SumOfKernels.scala:15-65."""

    def __init__(self, kernel1: Kernel, kernel2: Kernel):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def get_hyperparameters(self):
        return np.concatenate([self.kernel1.get_hyperparameters(), self.kernel2.get_hyperparameters()])

    def set_hyperparameters(self, value):           # :22-26
        value = np.asarray(value, dtype=np.float64)
        n1 = self.kernel1.number_of_hyperparameters()
        self.kernel1.set_hyperparameters(value[:n1])
        self.kernel2.set_hyperparameters(value[n1:])
        return self

    def number_of_hyperparameters(self):
        return self.kernel1.number_of_hyperparameters() + self.kernel2.number_of_hyperparameters()

    def hyperparameter_boundaries(self):
        l1, u1 = self.kernel1.hyperparameter_boundaries()
        l2, u2 = self.kernel2.hyperparameter_boundaries()
        return np.concatenate([l1, l2]), np.concatenate([u1, u2])

    def get_training_vectors(self): return self.kernel1.get_training_vectors()

    def set_training_vectors(self, vectors):
        self.kernel1.set_training_vectors(vectors)
        self.kernel2.set_training_vectors(vectors)
        return self

    def training_kernel(self): return self.kernel1.training_kernel() + self.kernel2.training_kernel()

    def training_kernel_diag(self):
        return self.kernel1.training_kernel_diag() + self.kernel2.training_kernel_diag()

    def training_kernel_and_derivative(self):       # :50-55
        k1, d1 = self.kernel1.training_kernel_and_derivative()
        k2, d2 = self.kernel2.training_kernel_and_derivative()
        return k1 + k2, list(d1) + list(d2)

    def cross_kernel(self, test):                   # :57-58
        return self.kernel1.cross_kernel(test) + self.kernel2.cross_kernel(test)

    def self_kernel(self, test): return self.kernel1.self_kernel(test) + self.kernel2.self_kernel(test)

    @property
    def white_noise_var(self): return self.kernel1.white_noise_var + self.kernel2.white_noise_var

    def __str__(self):
        return " + ".join(s for s in (str(self.kernel1), str(self.kernel2)) if len(s) > 0)
