import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
n, d, n_e = 200_000, 16, 100
rng = np.random.default_rng(3)
X = rng.random((n, d), dtype=np.float32)
y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(n)
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-2) * sg.EyeKernel()
eng = sg.ProjectedProcessEngine(0)
eng.experts_upload_grouped(X, y, n_e)
for _ in range(2):
    eng.bcm_nll(k)
eng.close()
