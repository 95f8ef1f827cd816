import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
for (n, d, m) in [(1353, 5, 1000), (20000, 16, 1000), (300000, 16, 1000), (4096, 4, 256), (4096, 4, 384)]:
    rng = np.random.default_rng(1)
    X = rng.random((n, d), dtype=np.float32); y = rng.random(n)
    Z = X[:m].astype(np.float64)
    k = 1 * sg.ARDRBFKernel(np.full(d, 1.0)) + sg.const(1) * sg.EyeKernel()
    e = sg.ProjectedProcessEngine(0)
    e.set_precision(N.SGP_PREC_I8)
    try:
        e.begin(k, Z); e.accumulate(X, y); G, b = e.finish()
        e.set_precision(N.SGP_PREC_F64)
        e.begin(k, Z); e.accumulate(X, y); G2, b2 = e.finish()
        print(n, d, m, "ok dG=%.2e db=%.2e" % (np.abs(G-G2).max()/np.abs(G2).max(), np.abs(b-b2).max()/np.abs(b2).max()), flush=True)
    except Exception as ex:
        print(n, d, m, "FAIL", str(ex)[-400:], flush=True)
    try:
        e.close()
    except Exception:
        pass
