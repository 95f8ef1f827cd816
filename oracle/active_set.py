"""CPU restatement (test infrastructure) of commons/ActiveSetProvider.scala:58-139, GreedilyOptimizingActiveSetProvider
("Fast Forward Selection to Speed Up Sparse Gaussian Process Regression", Seeger et al. 2003, as the reference codes it).

Parity notes: the FIRST point comes from Spark's `takeSample(…, 1, seed)` (ASP:70), whose RNG stream is
unpinned -- here it is an explicit index.  Everything after it is deterministic arithmetic and is restated verbatim,
including the quirks:
  * `getNext` is handed `kernelInstance.whiteNoiseVar` as its `sigma2` (ASP:76): with the default kernel that is
    1 + sigma2, not the sigma2 parameter;
  * per expert the candidates are folded left to right with `(max(delta, oldMax), if (oldMax > delta) oldIdx else i)`
    (ASP:108-127): on ties the LATER point wins, and one NaN delta poisons the expert's maximum (math.max(NaN, x) = NaN),
    after which the expert is dropped by `.filter(!_._1.isNaN)` (ASP:131);
  * across experts `max()` keeps the FIRST of equal scores (Ordering.max: `if (gteq(x, y)) x else y`);
  * points already in the active set are not excluded.
"""
from __future__ import annotations

import math

import numpy as np

from .ppa import assert_sym_positive_definite


def _fold_expert(deltas: np.ndarray):
    """ASP:108-127: returns (maxDelta, maxIndex) of one expert's candidates."""
    old_max, old_idx = -1.7976931348623157e308, -1          # Double.MinValue
    for i, delta in enumerate(deltas):
        new_idx = old_idx if old_max > delta else i         # a NaN on either side compares false -> i
        old_max = float("nan") if (math.isnan(delta) or math.isnan(old_max)) else max(delta, old_max)   # math.max
        old_idx = new_idx
    return old_max, old_idx


def candidate_deltas(cross: np.ndarray, y: np.ndarray, kii: np.ndarray, k_inv: np.ndarray, pdm_inv: np.ndarray,
                     magic_vector: np.ndarray, sigma2: float) -> np.ndarray:
    """ASP:109-124 for all points of an expert at once.  cross: m x n_e (column i = k(activeSet, x_i))."""
    with np.errstate(all="ignore"):
        p = np.einsum("ji,jk,ki->i", cross, k_inv, cross)
        q = np.einsum("ji,jk,ki->i", cross, pdm_inv, cross)
        mu = cross.T @ magic_vector
        sigma = math.sqrt(sigma2)
        li = np.sqrt(kii - p)
        ksi = 1.0 / ((sigma / li) ** 2 + 1.0 - q)
        kappa = ksi * (1.0 + 2.0 * (sigma / li) ** 2)
        return -np.log(sigma / li) - (np.log(ksi) + ksi * (1.0 - kappa) / sigma2 * (y - mu) ** 2 - kappa + 2.0) / 2.0


def get_next(kmm: np.ndarray, experts, active_set: np.ndarray, sigma2: float):
    """ASP:83-137.  experts: list of (y_e, kernel_e with training vectors X_e and hyper-parameters set)."""
    k_inv = np.linalg.inv(kmm)
    crosses = [(y, k.cross_kernel(active_set), k) for y, k in experts]
    g = sum(c @ c.T for _, c, _ in crosses)
    b = sum(c @ y for y, c, _ in crosses)
    pdm = sigma2 * kmm + g
    assert_sym_positive_definite(pdm)
    pdm_inv = np.linalg.inv(pdm)
    magic_vector = np.linalg.solve(pdm, b)
    best = None
    for y, c, k in crosses:
        d = candidate_deltas(c, y, k.training_kernel_diag(), k_inv, pdm_inv, magic_vector, sigma2)
        md, mi = _fold_expert(d)
        if math.isnan(md):
            continue
        if best is None or not (best[0] >= md):              # Ordering.max keeps the earlier of equal scores
            best = (md, k.get_training_vectors()[mi])
    if best is None:
        raise ValueError("empty.max")                        # what RDD.max() throws on an empty RDD
    return best[1]


def greedy_active_set(active_set_size: int, experts, kernel_factory, hyperparameters, first_point) -> np.ndarray:
    """ASP:63-81 with the takeSample'd first point given explicitly."""
    active = np.asarray(first_point, dtype=np.float64).reshape(1, -1)
    theta = np.asarray(hyperparameters, dtype=np.float64)
    while len(active) < active_set_size:
        inst = kernel_factory().set_hyperparameters(theta).set_training_vectors(active)
        nxt = get_next(inst.training_kernel(), experts, active, inst.white_noise_var)
        active = np.vstack([active, nxt])
    return active
