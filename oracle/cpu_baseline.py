"""CPU baseline harness (TEST / BENCH INFRASTRUCTURE -- see oracle/__init__.py).

Times the oracle's restatement of `getMatrixKmnKnmAndVectorKmny` (PGPH:20-36) with the reference's own
structure -- experts of n_e points, per expert an m x n_e cross kernel, a FULL dgemm K_mn K_mn^T and a
dgemv -- on all host cores.  Parallelism mirrors Spark `local[P]`: P workers each fold their share of the
experts into a private (G, b) (the treeAggregate seqOp, PGPH:26-30) and the partials are summed (combOp,
PGPH:31-35).  Workers are forked processes (one per usable core, BLAS single-threaded inside a worker like
netlib-java's F2J) writing their partial into shared memory, so neither the GIL nor pickling of m x m
partials is in the timed region's way.  This is optimistic for the reference (native BLAS, vectorised
element evaluation, no JVM allocation/GC, no task serialisation of the m x m zero value)."""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np

from .ppa import group_for_experts, get_matrix_kmn_knm_and_vector_kmny


def usable_cores() -> int:
    """Cores this process may really use: the affinity mask, capped by the cgroup CPU quota (a container that sees
    128 logical CPUs but has a 16-CPU quota thrashes with 128 workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:                                     # pragma: no cover
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def _worker(w, workers, shm_name, m, X, y, Z, kernel_factory, theta, n_e, start_evt, done_q):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:                                     # pragma: no cover
        pass
    groups = group_for_experts(len(X), n_e)[w::workers]                      # GPC:26-31
    experts = [(y[idx], kernel_factory().set_training_vectors(X[idx]).set_hyperparameters(theta))
               for idx in groups]                                                # GPC:35-36, GPR:50
    shm = shared_memory.SharedMemory(name=shm_name)
    out = np.ndarray((workers, m * m + m), dtype=np.float64, buffer=shm.buf)
    done_q.put(("ready", w))
    start_evt.wait()
    G, b = get_matrix_kmn_knm_and_vector_kmny(experts, Z)
    out[w, :m * m] = G.ravel()
    out[w, m * m:] = b
    done_q.put(("done", w))
    shm.close()


def stats_parallel(X, y, Z, kernel_factory, theta, n_e: int = 100, workers: int | None = None):
    """Returns (G, b, seconds): seconds covers the per-expert work of all workers plus the final sum of the
    partials (expert grouping and process start-up are outside, as Spark's would be)."""
    workers = workers or usable_cores()
    n_experts = int(np.floor(len(X) / n_e + 0.5))
    workers = max(1, min(workers, n_experts))
    m = len(Z)
    shm = shared_memory.SharedMemory(create=True, size=workers * (m * m + m) * 8)
    try:
        ctx = mp.get_context("fork")
        start_evt, q = ctx.Event(), ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(w, workers, shm.name, m, X, y, Z, kernel_factory, theta, n_e,
                                                   start_evt, q)) for w in range(workers)]
        for p in procs:
            p.start()
        for _ in range(workers):
            assert q.get(timeout=600)[0] == "ready"
        t0 = time.perf_counter()
        start_evt.set()
        for _ in range(workers):
            assert q.get(timeout=3600)[0] == "done"
        out = np.ndarray((workers, m * m + m), dtype=np.float64, buffer=shm.buf)
        tot = out.sum(axis=0)
        dt = time.perf_counter() - t0
        for p in procs:
            p.join()
        return tot[:m * m].reshape(m, m).copy(), tot[m * m:].copy(), dt
    finally:
        shm.close()
        shm.unlink()
