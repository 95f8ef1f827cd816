"""CPU baseline harness (TEST / BENCH INFRASTRUCTURE -- see oracle/__init__.py).

Times the oracle's restatement of `getMatrixKmnKnmAndVectorKmny` (PGPH:20-36) with the reference's own
structure -- experts of n_e points, per expert an m x n_e cross kernel, a FULL dgemm K_mn K_mn^T and a
dgemv -- on the host cores.  Parallelism mirrors Spark `local[P]`: P worker threads each fold their share
of the experts into a private (G, b) (the treeAggregate seqOp, PGPH:26-30) and the partials are summed
(combOp, PGPH:31-35); BLAS is single-threaded inside a worker like netlib-java's F2J.  NumPy releases the
GIL inside the heavy array ops, so the threads do run concurrently.  This is optimistic for the reference
(native BLAS, no JVM allocation/GC, no task serialisation)."""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .ppa import get_expert_labels_and_kernels, get_matrix_kmn_knm_and_vector_kmny


def stats_parallel(X, y, Z, kernel_factory, theta, n_e: int = 100, workers: int | None = None):
    workers = workers or os.cpu_count() or 1
    experts = get_expert_labels_and_kernels(X, y, kernel_factory, n_e)
    for _, k in experts:
        k.set_hyperparameters(theta)
    parts = [experts[w::workers] for w in range(workers)]
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1, user_api="blas")
    except Exception:                                     # pragma: no cover
        import contextlib
        ctx = contextlib.nullcontext()
    t0 = time.perf_counter()
    with ctx:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            res = list(ex.map(lambda p: get_matrix_kmn_knm_and_vector_kmny(p, Z), parts))
    G = sum(r[0] for r in res)
    b = sum(r[1] for r in res)
    return G, b, time.perf_counter() - t0
