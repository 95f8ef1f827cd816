import sys, time, cProfile, pstats, io
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
import bench
X, y = bench.make_shard("configs1", 0)
d, m = 16, 1000
gp = (sg.GaussianProcessRegression().setKernel(lambda: 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(18.0 / d))) + sg.const(1) * sg.EyeKernel())
      .setDatasetSizeForExpert(100).setActiveSetSize(m).setSigma2(bench.SIGMA2).setMaxIter(10).setTol(1e-6).setSeed(13))
gp.fit(X[:50000], y[:50000])
t0 = time.perf_counter(); e = sg.ProjectedProcessEngine(0); t1 = time.perf_counter(); e.close(); print("engine create %.1f ms" % (1e3*(t1-t0)))
for rep in range(2):
    t0 = time.perf_counter(); gp.fit(X, y); print("fit %.1f ms" % (1e3*(time.perf_counter()-t0)))
pr = cProfile.Profile(); pr.enable(); gp.fit(X, y); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
