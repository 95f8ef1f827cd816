"""One int8 statistics pass (device-resident shard) -- the target of ncu captures.
    python tools/prof_i8.py [i8|i8d|f64] [n] [d] [m]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N
mode = {"i8": N.SGP_PREC_I8, "f64": N.SGP_PREC_F64, "i8d": N.SGP_PREC_I8_DIRECT}[sys.argv[1] if len(sys.argv) > 1 else "i8"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 16
m = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
rng = np.random.default_rng(1)
X = rng.random((n, d), dtype=np.float32); y = rng.random(n)
Z = X[:m].astype(np.float64)
k = 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel() + sg.const(1e-4) * sg.EyeKernel()
e = sg.ProjectedProcessEngine(0)
e.set_precision(mode)
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y).cuda()
for rep in range(2):
    e.begin(k, Z); e.accumulate_ptr(Xd.data_ptr(), True, yd.data_ptr(), n, device=True); e.finish(copy_out=False)
    ms, nl = e.gram_kernel_time()
    print("rep %d: gram kernels %.3f ms over %d launches -> %.1f Mpts/s" % (rep, ms, nl, n / ms / 1e3), flush=True)
