"""Host-side mirror of the reference's kernel DSL / Params (no GPU needed): hyperparameter layout,
bounds, flattening into the C-ABI term list -- checked against the oracle's restatement."""
import numpy as np
import pytest

import oracle
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N


def _pairs():
    return [
        (lambda: 1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel(),
         lambda: 1 * oracle.ARDRBFKernel(5) + oracle.const(1) * oracle.EyeKernel()),
        (lambda: sg.Scalar(1.0).between(0).and_(30) * sg.RBFKernel(0.1, 1e-6, 10) + sg.WhiteNoiseKernel(0.5, 0, 1),
         lambda: oracle.Scalar(1.0).between(0).and_(30) * oracle.RBFKernel(0.1, 1e-6, 10) + oracle.WhiteNoiseKernel(0.5, 0, 1)),
        (lambda: sg.RBFKernel(10), lambda: oracle.RBFKernel(10)),
        (lambda: 2.0 * (1.5 * sg.ARDRBFKernel(np.array([0.2, 0.3])) + sg.const(0.5) * sg.EyeKernel()),
         lambda: 2.0 * (1.5 * oracle.ARDRBFKernel(np.array([0.2, 0.3])) + oracle.const(0.5) * oracle.EyeKernel())),
    ]


@pytest.mark.parametrize("idx", range(4))
def test_dsl_matches_oracle(idx):
    mk, mo = _pairs()[idx]
    k, o = mk(), mo()
    assert np.allclose(k.getHyperparameters(), o.get_hyperparameters())
    assert k.numberOfHyperparameters() == o.number_of_hyperparameters()
    for a, b in zip(k.hyperparameterBoundaries(), o.hyperparameter_boundaries()):
        assert np.array_equal(a, b)
    assert np.isclose(k.whiteNoiseVar, o.white_noise_var)
    theta = k.getHyperparameters() * 1.7 + 0.01
    k.setHyperparameters(theta); o.set_hyperparameters(theta)
    assert np.allclose(k.getHyperparameters(), o.get_hyperparameters())
    assert np.isclose(k.whiteNoiseVar, o.white_noise_var)
    assert str(k) == str(o)
    # self kernel = sum of every leaf's scale (each leaf has k(x,x)=1)
    assert np.isclose(sum(t["scale"] for t in k.flatten()), o.self_kernel(np.zeros(5)))


def test_flatten_scales_multiply_through_nesting():
    k = 2.0 * (1.5 * sg.ARDRBFKernel(np.array([0.2, 0.3])) + sg.const(0.5) * sg.EyeKernel())
    t = k.flatten()
    assert [x["type"] for x in t] == [N.SGP_TERM_ARD, N.SGP_TERM_EYE]
    assert np.isclose(t[0]["scale"], 3.0) and np.isclose(t[1]["scale"], 1.0)
    assert np.allclose(k.getHyperparameters(), [2.0, 1.5, 0.2, 0.3])


def test_params_defaults_and_getkernel():
    gp = sg.GaussianProcessRegression()
    assert (gp._datasetSizeForExpert, gp._activeSetSize, gp._sigma2, gp._maxIter, gp._tol) == (100, 100, 1e-3, 100, 1e-6)
    k = gp.setSigma2(1e-4).setKernel(lambda: 1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel()).getKernel()
    assert np.isclose(k.whiteNoiseVar, 1.0001)            # GPC:18 appends sigma2.const * Eye
    assert len(k.getHyperparameters()) == 6
    with pytest.raises(ValueError):
        sg.Scalar(-1.0) * sg.RBFKernel()
    with pytest.raises(ValueError):
        sg.Scalar(1.0, 2.0, 1.0)


# ---- hyper-parameter descriptors (what sgp_bcm_nll consumes) vs the oracle's trainingKernelAndDerivative ------------
def _dense_from_descriptors(k, X):
    """K and dK/dtheta_i rebuilt in numpy from (flatten(), hyper_descriptors()) exactly as csrc/bcm_nll.cu does."""
    terms = k.flatten()
    n, d = X.shape
    kt, sq = [], []
    for t in terms:
        if t["type"] == N.SGP_TERM_EYE:
            kt.append(np.eye(n)); sq.append(np.zeros((n, n))); continue
        beta = t["beta"] if t["type"] == N.SGP_TERM_ARD else np.full(d, 1.0 / (np.sqrt(2.0) * t["sigma"]))
        diff = X[:, None, :] - X[None, :, :]
        kt.append(np.exp(-((diff * beta) ** 2).sum(2))); sq.append((diff ** 2).sum(2))
    K = sum(t["scale"] * m for t, m in zip(terms, kt))
    dK = []
    for h in k.hyper_descriptors():
        if h["kind"] == N.SGP_HYPER_SCALE:
            dK.append(sum(c * kt[t] for t, c in h["coef"].items()))
        elif h["kind"] == N.SGP_HYPER_ARD_BETA:
            t, kk = h["term"], h["dim"]
            dx2 = (X[:, None, kk] - X[None, :, kk]) ** 2
            dK.append(terms[t]["scale"] * (-2.0 * h["value"] * dx2) * kt[t])
        else:
            t = h["term"]
            dK.append(terms[t]["scale"] * sq[t] * kt[t] / h["value"] ** 3)
    return K, dK


@pytest.mark.parametrize("idx", range(5))
def test_hyper_descriptors_match_oracle_derivatives(idx):
    pairs = _pairs() + [
        (lambda: 0.7 * (1.5 * sg.ARDRBFKernel(np.full(5, 0.9)) + 0.5 * sg.RBFKernel(2.0)) + sg.WhiteNoiseKernel(0.4, 0, 1) + sg.const(0.2) * sg.EyeKernel(),
         lambda: 0.7 * (1.5 * oracle.ARDRBFKernel(np.full(5, 0.9)) + 0.5 * oracle.RBFKernel(2.0)) + oracle.WhiteNoiseKernel(0.4, 0, 1) + oracle.const(0.2) * oracle.EyeKernel())]
    mk, mo = pairs[idx]
    k, o = mk(), mo()
    d = 2 if idx == 3 else 5
    X = np.random.default_rng(idx).random((9, d))
    K, dK = _dense_from_descriptors(k, X)
    K0, dK0 = o.set_training_vectors(X).training_kernel_and_derivative()
    assert len(dK) == len(dK0) == k.numberOfHyperparameters()
    assert np.allclose(K, K0, rtol=1e-13, atol=1e-15)
    for a, b in zip(dK, dK0):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("idx", range(5))
def test_descriptor_gradient_from_pair_sums(idx):
    """csrc/expert_common.cuh: sum_ab dK_i[a,b] W_ab for a symmetric W from ONE lower-triangle sweep of pair sums
    (S_t, Q_t, D_tk per non-Eye term, trW) -- the algebra both objective kernels use -- vs the oracle's dense dK_i."""
    pairs = _pairs() + [
        (lambda: 0.7 * (1.5 * sg.ARDRBFKernel(np.full(5, 0.9)) + 0.5 * sg.RBFKernel(2.0)) + sg.WhiteNoiseKernel(0.4, 0, 1) + sg.const(0.2) * sg.EyeKernel(),
         lambda: 0.7 * (1.5 * oracle.ARDRBFKernel(np.full(5, 0.9)) + 0.5 * oracle.RBFKernel(2.0)) + oracle.WhiteNoiseKernel(0.4, 0, 1) + oracle.const(0.2) * oracle.EyeKernel())]
    mk, mo = pairs[idx]
    k, o = mk(), mo()
    d = 2 if idx == 3 else 5
    rng = np.random.default_rng(10 + idx)
    n = 11
    X = rng.random((n, d))
    A = rng.standard_normal((n, n)); W = A + A.T
    _, dK0 = o.set_training_vectors(X).training_kernel_and_derivative()
    terms = k.flatten()
    tri = [(a, b) for a in range(n) for b in range(a + 1)]
    w2 = {(a, b): (1.0 if a == b else 2.0) * W[a, b] for a, b in tri}
    S, Q, D = {}, {}, {}
    for t, term in enumerate(terms):
        if term["type"] == N.SGP_TERM_EYE:
            continue
        beta = term["beta"] if term["type"] == N.SGP_TERM_ARD else np.full(d, 1.0 / (np.sqrt(2.0) * term["sigma"]))
        S[t], Q[t], D[t] = 0.0, 0.0, np.zeros(d)
        for a, b in tri:
            dx = X[a] - X[b]
            kw = np.exp(-((dx * beta) ** 2).sum()) * w2[(a, b)]
            S[t] += kw; Q[t] += (dx ** 2).sum() * kw; D[t] += dx ** 2 * kw
    trW = np.trace(W)
    hd = k.hyper_descriptors()
    assert len(hd) == len(dK0)
    for h, dk in zip(hd, dK0):
        if h["kind"] == N.SGP_HYPER_SCALE:
            g = sum(c * (trW if terms[t]["type"] == N.SGP_TERM_EYE else S[t]) for t, c in h["coef"].items())
        elif h["kind"] == N.SGP_HYPER_ARD_BETA:
            g = terms[h["term"]]["scale"] * (-2.0 * h["value"]) * D[h["term"]][h["dim"]]
        else:
            g = terms[h["term"]]["scale"] * Q[h["term"]] / h["value"] ** 3
        g0 = float((dk * W).sum())
        assert abs(g - g0) <= 1e-12 * max(1.0, abs(g0))


def test_expert_packing_matches_reference_grouping():
    from spark_gp_b200.hyperopt import group_for_experts, pack_experts
    for n, ne in ((1503, 100), (150, 100), (1000, 100), (999, 37)):
        g1, g0 = group_for_experts(n, ne), oracle.group_for_experts(n, ne)
        assert len(g1) == len(g0) and all(np.array_equal(a, b) for a, b in zip(g1, g0))
    X = np.arange(30.0).reshape(15, 2); y = np.arange(15.0)
    Xp, yp, off = pack_experts(X, y, 5)                     # E = 3 experts: points i % 3
    assert list(off) == [0, 5, 10, 15]
    assert np.array_equal(yp[:5], [0, 3, 6, 9, 12]) and np.array_equal(Xp[5], X[1])
    with pytest.raises(ZeroDivisionError):
        group_for_experts(40, 100)


def test_greedy_provider_selection_matches_scalar_fold():
    """GreedilyOptimizingActiveSetProvider.select_index (vectorised) vs the oracle's transcription of the reference's
    per-expert foldLeft + filter(!isNaN) + max (ActiveSetProvider.scala:108-135)."""
    from oracle.active_set import _fold_expert
    rng = np.random.default_rng(2)
    for trial in range(200):
        n = int(rng.integers(5, 90)); E = int(rng.integers(1, 9))
        delta = np.round(rng.standard_normal(n), 1)                 # coarse values: plenty of ties
        if trial % 3 == 0:
            delta[rng.integers(n)] = np.nan
        best = None
        for e in range(E):
            md, mi = _fold_expert(delta[e::E])
            if np.isnan(md) or mi < 0:
                continue
            if best is None or not (best[0] >= md):
                best = (md, mi * E + e)
        if best is None:
            with pytest.raises(ValueError):
                sg.GreedilyOptimizingActiveSetProvider.select_index(delta, E)
        else:
            assert sg.GreedilyOptimizingActiveSetProvider.select_index(delta, E) == best[1]


def test_classification_model_surface_matches_reference_semantics():
    """predictRaw / probability / prediction of GaussianProcessClassificationModel (GPCls:136-162 + Spark's
    raw2prediction) through a stub predictor, against the oracle's restatement."""
    from oracle.classification import classification_model_outputs
    from spark_gp_b200.classification import GaussianProcessClassificationModel

    class _Engine:
        def __init__(self, f): self.f = f
        def predict(self, X, with_variance=True): return self.f[:len(X)], None

    class _Raw:
        def __init__(self, f): self._engine = _Engine(f)

    f = np.array([-3.0, -0.2, 0.0, 0.4, 5.0])
    model = GaussianProcessClassificationModel(_Raw(f), np.zeros(1))
    raw0, prob0, pred0 = classification_model_outputs(f)
    X = np.zeros((len(f), 2))
    assert np.array_equal(model.predictRaw(X), raw0)
    assert np.allclose(model.predictProbability(X), prob0, rtol=0, atol=1e-16)
    assert np.array_equal(model.predict(X), pred0)
    assert list(pred0) == [0.0, 0.0, 0.0, 1.0, 1.0]                      # f == 0: argmax takes the first maximum
    assert prob0[3, 0] > 0.5 and pred0[3] == 1.0                        # the quirk: P(class 0) = sigmoid(f) > 1/2, label 1


def test_kmeans_active_set_provider():
    """ActiveSetProvider.scala:22-46: centroids of K-means on the features (host-side library call, as in the reference)."""
    rng = np.random.default_rng(0)
    centres = rng.random((5, 3)) * 10
    X = np.concatenate([c + 0.05 * rng.standard_normal((40, 3)) for c in centres])
    prov = sg.KMeansActiveSetProvider(maxIter=20)
    A = prov(5, X, None, None, None, 13)
    assert A.shape == (5, 3) and A.dtype == np.float64 and A.flags["C_CONTIGUOUS"]
    assert np.array_equal(A, prov(5, X, None, None, None, 13))                    # same seed, same centroids
    d = np.linalg.norm(A[:, None, :] - centres[None, :, :], axis=2)
    assert np.all(d.min(axis=0) < 0.05)                                           # every true centre is recovered
    assert sg.KMeansActiveSetProvider().maxIter == 20                             # the reference's default


def test_scale_matches_oracle():
    X = np.random.default_rng(1).random((50, 4)) * [1.0, 10.0, 0.0, 3.0] + [0.0, 5.0, 2.0, -1.0]   # column 2 is constant
    got, want = sg.scale(X), oracle.scale(X)
    assert np.array_equal(got, want)
    assert np.allclose(got.mean(0), 0.0, atol=1e-14) and np.allclose(got[:, [0, 1, 3]].std(0), 1.0) and np.all(got[:, 2] == 0.0)


def test_greedy_rank1_update_identities():
    """The algebra of csrc/greedy.cu restated in NumPy and checked against the reference's per-round quantities
    (oracle.active_set: inv(K_mm), inv(s2 K_mm + G), magic vector; p_i, q_i, mu_i of ASP:109-113): adding one active point
    borders K_mm and A = s2 K_mm + G, so
        p_i' = p_i + (u~.k_i - k*_i)^2 / s~,   q_i' = q_i + (u.k_i - k*_i)^2 / s,   mu_i' = mu_i + a (u.k_i - k*_i)
    with u~ = inv(K_mm) c, s~ = kii - c.u~, u = inv(A) w, w = s2 c + K_mn k*, s = s2 kii + k*.k* - w.u, a = (u.b - k*.y) / s."""
    rng = np.random.default_rng(9)
    n, d, rounds = 400, 3, 12
    X = rng.random((n, d)); y = np.sin(3 * X.sum(1)) + 0.1 * rng.standard_normal(n)
    beta = np.full(d, 2.5)
    kern = lambda: 1.2 * oracle.ARDRBFKernel(beta) + oracle.const(0.3) * oracle.EyeKernel()
    k0 = kern()
    s2 = k0.white_noise_var
    kii = 1.2 + 0.3                                               # trainingKernelDiag: every leaf has k(x, x) = 1
    picks = rng.permutation(n)[:rounds]
    Kt = np.zeros((0, n)); Kinv = np.zeros((0, 0)); Ainv = np.zeros((0, 0)); b = np.zeros(0); mv = np.zeros(0)
    p = np.zeros(n); q = np.zeros(n); mu = np.zeros(n)
    for r, idx in enumerate(picks):
        kstar = kern().set_training_vectors(X[[idx]]).cross_kernel(X)[:, 0]      # k(x_i, x_idx) for every point i
        m = len(Kt)
        c = Kt[:, idx] if m else np.zeros(0)
        g = Kt @ kstar if m else np.zeros(0)
        gamma, bnew = kstar @ kstar, kstar @ y
        ut = Kinv @ c if m else np.zeros(0)
        w = s2 * c + g
        u = Ainv @ w if m else np.zeros(0)
        s1 = kii - c @ ut
        s = s2 * kii + gamma - w @ u
        a = ((u @ b) - bnew) / s
        t1 = Kt.T @ ut if m else np.zeros(n)
        t2 = Kt.T @ u if m else np.zeros(n)
        p += (t1 - kstar) ** 2 / s1; q += (t2 - kstar) ** 2 / s; mu += a * (t2 - kstar)
        # bordered inverses
        def border(inv, v, sc):
            out = np.zeros((m + 1, m + 1))
            out[:m, :m] = inv + np.outer(v, v) / sc
            out[:m, m] = out[m, :m] = -v / sc
            out[m, m] = 1.0 / sc
            return out
        Kinv, Ainv = border(Kinv, ut, s1), border(Ainv, u, s)
        mv = np.concatenate([mv + u * a, [-a]]); b = np.concatenate([b, [bnew]])
        Kt = np.vstack([Kt, kstar])
        # the reference's quantities for this active set, from scratch
        active = X[picks[:r + 1]]
        inst = kern().set_training_vectors(active)
        kmm = inst.training_kernel()
        cross = inst.cross_kernel(X).T                            # m x n (crossKernel: rows = test vectors, Kernel.scala:69-74)
        G = cross @ cross.T
        A = s2 * kmm + G
        assert np.allclose(Kinv, np.linalg.inv(kmm), rtol=1e-8, atol=1e-10)
        assert np.allclose(Ainv, np.linalg.inv(A), rtol=1e-7, atol=1e-12)
        assert np.allclose(mv, np.linalg.solve(A, cross @ y), rtol=1e-7, atol=1e-12)
        assert np.allclose(p, np.einsum("ji,jk,ki->i", cross, np.linalg.inv(kmm), cross), rtol=1e-8, atol=1e-12)
        assert np.allclose(q, np.einsum("ji,jk,ki->i", cross, np.linalg.inv(A), cross), rtol=1e-7, atol=1e-12)
        assert np.allclose(mu, cross.T @ np.linalg.solve(A, cross @ y), rtol=1e-7, atol=1e-12)


def test_bcm_register_kernel_identities():
    """The two identities `bcm_nll_reg_kernel` (csrc/bcm_nll.cu) rests on, in NumPy: (1) n sweeps of Goodnight's sweep
    operator turn an SPD matrix into minus its inverse, the pivots are the Schur complements (all positive) and their logs
    sum to log|det|; (2) the per-dimension gradient sums are quadratic forms,
    sum_ab (x_ak - x_bk)^2 M_ab = 2 (sum_a x_ak^2 m_a - x_k' M x_k) with m = M 1, for any symmetric M."""
    rng = np.random.default_rng(4)
    n, d = 37, 5
    X = rng.standard_normal((n, d))
    K = np.exp(-0.5 * ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)) + 0.1 * np.eye(n)
    A = K.copy()
    logdet = 0.0
    for k in range(n):
        piv = A[k, k]
        assert piv > 0
        logdet += np.log(piv)
        col = A[:, k].copy()
        A -= np.outer(col, col) / piv                      # rank-1 update everywhere ...
        A[:, k] = col / piv; A[k, :] = col / piv            # ... then the pivot column / row
        A[k, k] = -1.0 / piv
    assert np.allclose(-A, np.linalg.inv(K), rtol=1e-9, atol=1e-10)
    assert abs(logdet - np.linalg.slogdet(K)[1]) < 1e-9
    W = rng.standard_normal((n, n)); M = K * (W + W.T)      # any symmetric pair weight
    m1 = M.sum(1)
    for k in range(d):
        direct = (((X[:, None, k] - X[None, :, k]) ** 2) * M).sum()
        assert abs(direct - 2.0 * ((X[:, k] ** 2) @ m1 - X[:, k] @ M @ X[:, k])) < 1e-9 * max(1.0, abs(direct))
