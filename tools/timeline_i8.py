"""In-kernel clock64 timeline of the int8 Gram kernel (debug instantiation): per-role event gaps, units 64..95, for the
publishing diagonal CTA (0,0) and the consuming CTA (1,0).   python tools/timeline_i8.py [n] [d] [m]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rng = np.random.default_rng(13)
X = rng.random((n, d), dtype=np.float32)
y = rng.random(n)
Z = X[:m].astype(np.float64)
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel()
import os; os.environ.setdefault("SGP_I8_IMPL", "ring")
eng = sg.ProjectedProcessEngine(0)
eng.set_precision(N.SGP_PREC_I8)
eng.debug_i8_tile()                       # arm
import torch
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y).cuda()
for _ in range(2):
    eng.begin(k, Z)
    eng.accumulate_ptr(Xd.data_ptr(), True, yd.data_ptr(), n, device=True)
    eng.finish(copy_out=False)
print("gram kernel ms (debug build):", eng.gram_kernel_time())
tl = eng.debug_i8_timeline()              # [cta][role][unit][event]
names = {0: ["wait_x", "x_ok", "q_ok", "issued"], 1: ["wait_pi", "pi_ok", "pj_ok", "issued"],
         2: ["q_wait", "q_ok", "ld", "exp", "pe_ok", "stored", "next_top"], 3: ["q_wait", "q_ok", "ld", "exp", "pe_ok", "stored", "next_top"],
         4: ["start", "a", "b", "c", "d", "e", "f"]}
roles = ["dist", "gram", "epi0", "epi1", "share"]
for cta, cname in enumerate(["publisher (0,0)", "consumer (1,0)"]):
    print("==", cname)
    t = tl[cta]
    for r in range(5):
        ev = t[r]
        units = [u for u in range(32) if ev[u, 0] > 0]
        if len(units) < 3:
            continue
        first = ev[units, 0]
        period = np.diff(first) / np.diff(units)
        line = "  %-5s period %6.0f |" % (roles[r], period.mean())
        nev = len(names[r])
        for e in range(1, nev):
            ok = [u for u in units if ev[u, e] > 0 and ev[u, e - 1] > 0]
            if ok:
                line += " %s-%s %5.0f" % (names[r][e - 1], names[r][e], np.mean(ev[ok, e] - ev[ok, e - 1]))
        print(line)
    base = t[t > 0].min()
    print("  window start (clk since first event of either CTA): %d" % (base - tl[tl > 0].min()))
    g0 = tl[tl > 0].min()
    for u in (10, 11, 12):
        print("  unit +%d:" % u, {roles[r]: [int(x - g0) if x > 0 else -1 for x in t[r, u, :7]] for r in range(5)})

P = eng.i8_progress if hasattr(eng, "i8_progress") else None
if P is not None:
    g0 = P[P > 0].min()
    print("per-CTA progress: clk/unit between marks (every 128 units), CTA = slice*36 + tile")
    for cta in list(range(0, 10)) + [35, 36, 37, 72, 108, 143]:
        t = P[cta]
        k = int((t > 0).sum())
        if k < 2:
            continue
        rate = np.diff(t[:k]) / 128.0
        print("  cta %3d start %8d  rates:" % (cta, t[0] - g0), " ".join("%4.0f" % r for r in rate))
eng.close()
