import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
e = sg.ProjectedProcessEngine(0)
rng = np.random.default_rng(13)
d, m = 16, 1000
for n in (32768, 300_000):
    for width in (24.0, 30.0, 33.0):
        for s2 in (1e-4, 1e-2):
            X = rng.random((n, d), dtype=np.float32)
            y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(n)
            Z = X[rng.permutation(n)[:m]].astype(np.float64)
            k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(width / d))) + sg.const(1) * sg.EyeKernel() + sg.const(s2) * sg.EyeKernel()
            Xt = rng.random((1000, d))
            out = {}
            for mode in (N.SGP_PREC_F64_STRICT, N.SGP_PREC_AUTO, N.SGP_PREC_F64):
                e.set_precision(mode); e.begin(k, Z); e.accumulate(X, y); G, b = e.finish(); path = e.last_path(); e.magic(); out[mode] = (G,) + e.predict(Xt) + (path,)
            r = out[N.SGP_PREC_F64_STRICT]
            line = "N=%7d beta^2 d=%g sigma2=%g" % (n, width, s2)
            for mode, nm in ((N.SGP_PREC_AUTO, "auto"), (N.SGP_PREC_F64, "f64")):
                o = out[mode]
                line += "  %s(path %d): dmean=%.1e dvar=%.1e" % (nm, o[3], np.abs(o[1] - r[1]).max() / np.abs(r[1]).max(), np.abs(o[2] / r[2] - 1).max())
            print(line, flush=True)
