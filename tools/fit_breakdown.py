"""Where the whole-fit time goes (bench.py's fit number): expert upload, one BCM objective evaluation, statistics, tail.
    python tools/fit_breakdown.py [n] [d] [m] [n_e]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
n_e = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rng = np.random.default_rng(3)
X = rng.random((n, d), dtype=np.float32)
y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(n)
Z = X[:m].astype(np.float64)
kern = lambda: 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-2) * sg.EyeKernel()
eng = sg.ProjectedProcessEngine(0)


def timed(label, fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); eng.sync(); ts.append(time.perf_counter() - t0)
    print("%-34s %8.2f ms (min of %d)" % (label, 1e3 * min(ts), reps), flush=True)
    return r


timed("experts_upload_grouped", lambda: eng.experts_upload_grouped(X, y, n_e))
k = kern()
timed("bcm_nll (value + gradient)", lambda: eng.bcm_nll(k), reps=5)
timed("statistics (host X, AUTO)", lambda: eng.statistics(k, Z, X, y))
timed("magic (tail)", lambda: eng.magic(copy_out=False))
print("last bcm path:", eng.last_bcm_path())
eng.close()
