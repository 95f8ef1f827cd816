"""Prints the in-kernel clock64 timeline of one off-diagonal CTA of the I8 kernel (GPU box)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
n, d, m = 600_000, 16, 1000
rng = np.random.default_rng(1)
X = rng.random((n, d), dtype=np.float32); y = rng.random(n)
Z = X[:m].astype(np.float64)
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
e = sg.ProjectedProcessEngine(0)
e.set_precision(N.SGP_PREC_I8)
e.debug_i8_tile()
e.begin(k, Z); e.accumulate(X[:500_000], y[:500_000]); e.finish(copy_out=False)
tl = e.debug_i8_timeline()
t0 = tl[tl > 0].min()
tl = np.where(tl > 0, tl - t0, -1)
names = {0: ["dist:wait_x", "dist:x_ok", "dist:issued", "gram:wait_p", "gram:p_ok", "gram:issued"],
         1: ["I:start", "I:done", "J:start", "J:done"],
         2: ["I:start", "I:loaded", "I:pempty_ok", "I:stored", "J:start", "J:loaded", "J:pempty_ok", "J:stored"]}
for u in range(4, 10):
    print("unit %d" % (64 + u))
    for role, rn in ((0, "MMA "), (1, "EXP "), (2, "PACK")):
        print("   %s  " % rn + "  ".join("%s=%d" % (names[role][ev], tl[role, u, ev]) for ev in range(len(names[role]))))
per = np.diff(tl[2, 2:30, 7])
print("pack 'J stored' period per unit: mean %.0f  min %d max %d" % (per.mean(), per.min(), per.max()))
a = tl[1, 2:30]
print("EXP  mean: I tile %.0f  gap %.0f  J tile %.0f" % ((a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean()))
a = tl[2, 2:30]
print("PACK mean: I wait+load %.0f  pempty %.0f  pack+store %.0f | J wait+load %.0f pempty %.0f pack+store %.0f" % (
    (a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean(),
    (a[:, 5] - a[:, 4]).mean(), (a[:, 6] - a[:, 5]).mean(), (a[:, 7] - a[:, 6]).mean()))
a = tl[0, 2:30]
print("MMA mean: dist wait_x %.0f  dist issue(+q_empty waits) %.0f   gram wait_p %.0f  gram issue %.0f" % ((a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 4] - a[:, 3]).mean(), (a[:, 5] - a[:, 4]).mean()))
