"""GPU-side timing of the m x m tail pieces (run on the GPU box)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
for m in (1000, 2000):
    rng = np.random.default_rng(0)
    d = 16
    X = rng.random((50000, d), dtype=np.float32); y = rng.random(50000)
    Z = X[:m].astype(np.float64)
    k = 1 * sg.ARDRBFKernel(np.full(d, 1.06)) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
    e = sg.ProjectedProcessEngine(0)
    e.begin(k, Z); e.accumulate(X, y); e.finish(copy_out=False)
    for i in range(3):
        t0 = time.perf_counter(); e.magic(copy_out=False); print("m=%d magic call %d: %.1f ms" % (m, i, 1e3 * (time.perf_counter() - t0)), flush=True)
    e.close()
