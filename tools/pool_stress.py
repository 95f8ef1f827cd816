import sys
sys.path.insert(0, ".")
import numpy as np
import spark_gp_b200 as sg
def reg(n, d, m, terms=1):
    rng = np.random.default_rng(n + d + m)
    X = rng.random((n, d), dtype=np.float32); y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(n)
    if terms == 1:
        kf = lambda: 1 * sg.ARDRBFKernel(np.full(d, np.sqrt(12.0 / d))) + sg.const(1) * sg.EyeKernel()
    else:
        kf = lambda: 0.6 * sg.ARDRBFKernel(np.full(d, np.sqrt(8.0 / d))) + 0.5 * sg.RBFKernel(0.7) + sg.const(1) * sg.EyeKernel()
    gp = sg.GaussianProcessRegression().setKernel(kf).setDatasetSizeForExpert(100).setActiveSetSize(m).setSigma2(1e-2).setMaxIter(3).setSeed(1)
    model = gp.fit(X, y)
    pred = model.predict(X[:2000].astype(np.float64))
    rmse = float(np.sqrt(np.mean((pred - y[:2000]) ** 2)))
    return model, rmse, pred
rng = np.random.default_rng(0)
models = []
first = {}
for cfg in [(50_000, 8, 200, 1), (120_000, 16, 500, 1), (40_000, 5, 130, 2), (50_000, 8, 200, 1), (200_000, 40, 300, 1), (60_000, 8, 1000, 1)]:
    m, r, pred = reg(*cfg); models.append(m); print(cfg, "rmse %.4f" % r, "idle engines:", {k: len(v) for k, v in sg.engine.ProjectedProcessEngine._idle.items()}, flush=True)
    assert np.isfinite(r)
    if cfg in first:                                  # same data through recycled contexts: identical model
        assert np.array_equal(pred, first[cfg]), "pooled context changed the result"
        print("  identical to the first fit of this configuration")
    first[cfg] = pred
# earlier models still predict with their own contexts
Xt = rng.random((100, 8))
p0 = models[0].predict(Xt); p3 = models[3].predict(Xt)
assert np.all(np.isfinite(p0)) and np.all(np.isfinite(p3))
del models
import gc; gc.collect()
print("after release:", {k: len(v) for k, v in sg.engine.ProjectedProcessEngine._idle.items()})
# classification through the same pool
n = 6000; X = rng.standard_normal((n, 4)); yc = (X[:, 0] + 0.5 * X[:, 1] ** 2 > 0.3).astype(float)
gc_ = sg.GaussianProcessClassifier().setKernel(lambda: 1 * sg.RBFKernel(1.0)).setDatasetSizeForExpert(50).setActiveSetSize(100).setMaxIter(3)
cm = gc_.fit(X, yc)
acc = float(np.mean(cm.predict(X[:1000]) == yc[:1000])); print("classifier acc %.3f" % acc)
assert acc > 0.85
print("OK")
