"""Times one evaluation of the two hyper-parameter objectives at BASELINE configs[1] scale (GPU box):
10^4 experts of 100 points, d = 16 (regression NLL + gradient; classification Laplace)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200.hyperopt import pack_experts
n, d, n_e = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 16, 100
rng = np.random.default_rng(3)
X = rng.random((n, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
e = sg.ProjectedProcessEngine(0)
Xp, yp, off = pack_experts(X, y, n_e)
t0 = time.perf_counter(); e.experts_upload(Xp, yp, off); e.sync(); t_up = time.perf_counter() - t0
for rep in range(3):
    t0 = time.perf_counter(); v, g = e.bcm_nll(k); dt = time.perf_counter() - t0
    print("bcm_nll  rep %d: %.1f ms  (%d experts, %d hypers) -> %.2e points/s per evaluation   nll=%.6e" % (rep, dt * 1e3, len(off) - 1, len(g), n / dt, v), flush=True)
yc = (y > np.median(y)).astype(np.float64)
Xp, yp, off = pack_experts(X, yc, n_e)
e.experts_upload(Xp, yp, off)
for rep in range(3):
    t0 = time.perf_counter(); v, g = e.laplace_nll(k, 1e-6); dt = time.perf_counter() - t0
    print("laplace  rep %d: %.1f ms -> %.2e points/s per evaluation   nll=%.6e" % (rep, dt * 1e3, n / dt, v), flush=True)
print("upload %.1f ms" % (t_up * 1e3))
