"""N-GPU consistency check (GPU box; launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/multi_gpu_check.py

Every rank owns a contiguous share of the points (statistics) and of the experts (objectives).  The NCCL-reduced
results of `sgp_stats_finish`, `sgp_bcm_nll` and `sgp_laplace_nll` must equal what ONE context computes from all the
data (rank 0 recomputes that on its own GPU without a communicator).  Tolerances: statistics 1e-12 relative in strict
mode (the only difference is the order of the final sums), objectives 1e-11.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
from spark_gp_b200.hyperopt import group_for_experts


def pack_groups(X, y, groups):
    order = np.concatenate(groups)
    off = np.concatenate([[0], np.cumsum([len(g) for g in groups])])
    return np.ascontiguousarray(X[order]), np.ascontiguousarray(y[order]), off


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    eng = sg.ProjectedProcessEngine(local)
    ids = [sg.ProjectedProcessEngine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    if world > 1:
        eng.comm_init(ids[0], rank, world)

    rng = np.random.default_rng(5)
    n, d, m, n_e = 40000, 6, 200, 100
    X = rng.random((n, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
    Z = X[rng.permutation(n)[:m]].copy()
    beta = np.full(d, 1.3)
    kern = 1.5 * sg.ARDRBFKernel(beta) + sg.const(1) * sg.EyeKernel() + sg.const(1e-2) * sg.EyeKernel()
    ok = True

    # ---- statistics: points sharded over ranks, one all-reduce in finish -------------------------------------
    lo, hi = rank * n // world, (rank + 1) * n // world
    for mode, tol in ((N.SGP_PREC_F64_STRICT, 1e-12), (N.SGP_PREC_F64, 1e-6)):
        eng.set_precision(mode)
        eng.begin(kern, Z)
        eng.accumulate(X[lo:hi], y[lo:hi])
        G, b = eng.finish()
        if rank == 0:
            solo = sg.ProjectedProcessEngine(local)
            solo.set_precision(mode)
            solo.begin(kern, Z)
            solo.accumulate(X, y)
            G1, b1 = solo.finish()
            solo.close()
            eg, eb = rel(G, G1), rel(b, b1)
            print("stats mode %d: dG=%.2e db=%.2e (tol %.0e)" % (mode, eg, eb, tol), flush=True)
            ok &= eg <= tol and eb <= tol

    # ---- objectives: experts sharded over ranks, rows all-reduced ---------------------------------------------
    groups = group_for_experts(n, n_e)
    mine = groups[rank::world]
    Xp, yp, off = pack_groups(X, y, mine)
    eng.experts_upload(Xp, yp, off)
    v, g = eng.bcm_nll(kern)
    ycls = (y > np.median(y)).astype(np.float64)
    Xc, yc, offc = pack_groups(X[:8000], ycls[:8000], group_for_experts(8000, n_e)[rank::world])
    eng.experts_upload(Xc, yc, offc)
    vl, gl = eng.laplace_nll(kern, 1e-6)
    vl2, gl2 = eng.laplace_nll(kern, 1e-6)          # warm start
    if rank == 0:
        solo = sg.ProjectedProcessEngine(local)
        Xa, ya, offa = pack_groups(X, y, groups)
        solo.experts_upload(Xa, ya, offa)
        v1, g1 = solo.bcm_nll(kern)
        Xa, ya, offa = pack_groups(X[:8000], ycls[:8000], group_for_experts(8000, n_e))
        solo.experts_upload(Xa, ya, offa)
        w1, h1 = solo.laplace_nll(kern, 1e-6)
        w2, h2 = solo.laplace_nll(kern, 1e-6)
        solo.close()
        e = (abs(v - v1) / abs(v1), rel(g, g1), abs(vl - w1) / abs(w1), rel(gl, h1), abs(vl2 - w2) / abs(w2), rel(gl2, h2))
        print("bcm_nll: dval=%.2e dgrad=%.2e | laplace: dval=%.2e dgrad=%.2e | warm: dval=%.2e dgrad=%.2e" % e, flush=True)
        ok &= max(e) <= 1e-11
        print("MULTI_GPU_CHECK %s (world=%d)" % ("PASS" if ok else "FAIL", world), flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
