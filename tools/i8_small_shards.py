import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
d, m = 16, 1000
e = sg.ProjectedProcessEngine(0)
for seed in (13, 14):
    rng = np.random.default_rng(seed)
    NMAX = 262144
    X = rng.random((NMAX, d), dtype=np.float32)
    y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(NMAX)
    Z = X[rng.permutation(NMAX)[:m]].astype(np.float64)
    Xt = rng.random((1000, d))
    k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
    for n in (16384, 32768, 65536, 131072, 262144):
        out = {}
        for mode in (N.SGP_PREC_F64_STRICT, N.SGP_PREC_F64, N.SGP_PREC_I8, N.SGP_PREC_I8_DIRECT):
            e.set_precision(mode); e.begin(k, Z); e.accumulate(X[:n], y[:n])
            G, b = e.finish(); e.magic(); out[mode] = (G, ) + e.predict(Xt)
        r = out[N.SGP_PREC_F64_STRICT]
        line = "seed %d N=%7d" % (seed, n)
        for mode, nm in ((N.SGP_PREC_F64, "f64"), (N.SGP_PREC_I8, "i8"), (N.SGP_PREC_I8_DIRECT, "i8d")):
            o = out[mode]
            line += "  %s: dG=%.1e dmean=%.2e dvar=%.1e" % (nm, np.abs(o[0] - r[0]).max() / np.abs(r[0]).max(), np.abs(o[1] - r[1]).max() / np.abs(r[1]).max(), np.abs(o[2] / r[2] - 1).max())
        print(line, flush=True)
