import sys; sys.path.insert(0, ".")
import numpy as np, os
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
c = np.load("tests/golden/airfoil_case.npz")
kernel = (1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel() + sg.const(float(c["sigma2"])) * sg.EyeKernel())
kernel.setHyperparameters(c["theta"])
e = sg.ProjectedProcessEngine(0)
X, Z = c["X"], c["Z"]
s = np.sqrt(np.log2(np.e)) * c["theta"][1:]
xh = (X - Z.mean(0)) * s
print("scaled |x|^2: mean %.2f max %.2f ; |z|^2 max %.2f" % ((xh**2).sum(1).mean(), (xh**2).sum(1).max(), (((Z - Z.mean(0)) * s)**2).sum(1).max()))
for mode in (N.SGP_PREC_F64, N.SGP_PREC_I8):
    e.set_precision(mode); e.begin(kernel, Z); e.accumulate(X, c["y"]); G, b = e.finish()
    gmax = np.abs(c["G_diag"]).max()
    print("mode", mode, "diag %.2e row0 %.2e sum %.2e b %.2e" % (np.abs(np.diag(G) - c["G_diag"]).max() / gmax, np.abs(G[0] - c["G_row0"]).max() / gmax, abs(G.sum() - c["G_sum"]) / abs(c["G_sum"]), np.abs(b - c["b"]).max() / np.abs(c["b"]).max()))
    e.magic(); mean, var = e.predict(c["Xtest"])
    print("   dmean %.2e dvar %.2e" % (np.abs(mean - c["mean"]).max() / np.abs(c["mean"]).max(), np.abs(var / c["var"] - 1).max()))
