"""GPU: posterior-mean gap between the int8 kernel and the all-fp64 kernel as N grows (design evidence for AUTO's
size gate).  BASELINE-style data: U[0,1)^16, m=1000, theta as in bench.py."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
d, m = 16, 1000
rng = np.random.default_rng(13)
NMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
X = rng.random((NMAX, d), dtype=np.float32)
y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(NMAX)
Z = X[rng.permutation(NMAX)[:m]].astype(np.float64)
Xt = rng.random((1000, d))
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
e = sg.ProjectedProcessEngine(0)
n = 250_000
while n <= NMAX:
    out = {}
    for mode in (N.SGP_PREC_F64_STRICT, N.SGP_PREC_F64, N.SGP_PREC_I8):
        e.set_precision(mode); e.begin(k, Z)
        for s in range(0, n, 1_000_000):
            e.accumulate(X[s:min(n, s + 1_000_000)], y[s:min(n, s + 1_000_000)])
        G, b = e.finish(); e.magic(); out[mode] = (G, ) + e.predict(Xt)
    r = out[N.SGP_PREC_F64_STRICT]
    line = "N=%8d" % n
    for mode, nm in ((N.SGP_PREC_F64, "f64"), (N.SGP_PREC_I8, "i8")):
        o = out[mode]
        line += "  %s: dG=%.1e dmean=%.2e dvar=%.1e" % (nm, np.abs(o[0] - r[0]).max() / np.abs(r[0]).max(), np.abs(o[1] - r[1]).max() / np.abs(r[1]).max(), np.abs(o[2] / r[2] - 1).max())
    print(line, flush=True)
    n *= 2
