"""How the int8 Gram behaves on ill-conditioned problems: posterior mean / variance of the int8 (direct / tensor) and the
default fp64 kernel against the all-fp64 mode, with cond_2(A) and cond_1(A) of A = s2 K_mm + G -- the calibration data of
the condition gate in sgp_magic.  Cases: the benchmark data at several N and kernel widths, the airfoil fixture (standardised
features) replicated with jitter."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
e = sg.ProjectedProcessEngine(0)


def run(tag, kernel, X, y, Z, Xt, modes):
    out = {}
    for mode in (N.SGP_PREC_F64_STRICT,) + modes:
        e.set_precision(mode); e.begin(kernel, Z); e.accumulate(X, y)
        G, b = e.finish()
        try:
            e.magic(); out[mode] = (G, ) + e.predict(Xt)
        except Exception as ex:
            out[mode] = None
    r = out[N.SGP_PREC_F64_STRICT]
    s2 = float(kernel.whiteNoiseVar)
    A = s2 * (e.cross_kernel(Z) + s2 * np.eye(len(Z))) + r[0]
    line = "%-34s N=%7d cond2=%.1e cond1=%.1e" % (tag, len(X), np.linalg.cond(A), np.linalg.cond(A, 1))
    names = {N.SGP_PREC_F64: "f64", N.SGP_PREC_I8_DIRECT: "i8d", N.SGP_PREC_I8: "i8"}
    for mode in modes:
        o = out[mode]
        if o is None:
            line += "  %s: NOT PD" % names[mode]
        else:
            line += "  %s: dmean=%.1e dvar=%.1e" % (names[mode], np.abs(o[1] - r[1]).max() / np.abs(r[1]).max(), np.abs(o[2] / r[2] - 1).max())
    print(line, flush=True)


rng = np.random.default_rng(13)
d, m = 16, 1000
for n in (32768, 1_000_000):
    for width in (18.0, 6.0, 2.0):
        X = rng.random((n, d), dtype=np.float32)
        y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(n)
        Z = X[rng.permutation(n)[:m]].astype(np.float64)
        k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(width / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
        run("bench data, beta^2 d = %g" % width, k, X, y, Z, rng.random((1000, d)), (N.SGP_PREC_F64, N.SGP_PREC_I8, N.SGP_PREC_I8_DIRECT))
c = np.load(os.path.join("tests", "golden", "airfoil_case.npz"))
kernel = (1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel() + sg.const(float(c["sigma2"])) * sg.EyeKernel())
kernel.setHyperparameters(c["theta"])
for reps, jit in ((1, 0.0), (30, 0.05), (200, 0.05), (200, 0.3)):
    X = np.tile(c["X"], (reps, 1)); y = np.tile(c["y"], reps)
    if jit > 0:
        X = X + jit * rng.standard_normal(X.shape); y = y + 0.05 * rng.standard_normal(len(y))
    run("airfoil x %d, jitter %.2f" % (reps, jit), kernel, X, y, c["Z"], c["Xtest"], (N.SGP_PREC_F64, N.SGP_PREC_I8_DIRECT))
