"""Generates tests/golden/*.npz in THIS container (the GPU box has no /root/reference).

airfoil_case.npz -- BASELINE config 1: data/airfoil.csv of the reference, standardised as
commons/util/Scaling.scala:10-25 does, kernel `1*ARDRBFKernel(5) + 1.const*EyeKernel` + sigma2=1e-4
(regression/examples/Airfoil.scala:18-22), expert=100, active=1000; theta and the active set are explicit
(the reference's L-BFGS-B trajectory and takeSample are unpinned).  Outputs come from the fp64 oracle.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from oracle import ARDRBFKernel, EyeKernel, const

OUT = os.path.dirname(os.path.abspath(__file__))
REF_DATA = "/root/reference/data"


def airfoil():
    raw = np.loadtxt(os.path.join(REF_DATA, "airfoil.csv"), delimiter=",")
    X = oracle.scale(raw[:, :5])
    y = raw[:, 5].copy()
    rng = np.random.default_rng(20260921)
    perm = rng.permutation(len(X))
    test_idx, train_idx = np.sort(perm[:150]), np.sort(perm[150:])
    Xtr, ytr, Xte = X[train_idx], y[train_idx], X[test_idx]
    m = 1000
    active_idx = np.sort(rng.permutation(len(Xtr))[:m])
    Z = Xtr[active_idx]
    # a plausible optimum (hand-set; the point of the fixture is identical inputs, not the optimiser)
    theta = np.array([55.0, 0.9, 0.35, 1.4, 0.25, 1.1])
    user = lambda: 1 * ARDRBFKernel(5) + const(1) * EyeKernel()
    factory = oracle.get_kernel(user, 1e-4)
    experts = oracle.get_expert_labels_and_kernels(Xtr, ytr, factory, 100)
    pred, G, b = oracle.projected_process(experts, Z, factory, theta)
    mean, var = pred.predict_many(Xte)
    nll, grad = oracle.regression.bcm_objective(experts, theta)
    np.savez_compressed(os.path.join(OUT, "airfoil_case.npz"), X=Xtr, y=ytr, Xtest=Xte, Z=Z, theta=theta,
                        sigma2=1e-4, G_diag=np.diag(G).copy(), G_row0=G[0].copy(), G_sum=G.sum(), b=b,
                        magic_vector=pred.magic_vector, magic_matrix_diag=np.diag(pred.magic_matrix).copy(),
                        mean=mean, var=var, bcm_nll=nll, bcm_grad=grad, n_experts=len(experts))
    print("airfoil: N=%d m=%d  |mean|max=%.3f var range=(%.4f, %.4f) nll=%.6f" %
          (len(Xtr), m, np.abs(mean).max(), var.min(), var.max(), nll))


def small_synthetic():
    """Small seeded cases covering DSL shapes: multi-term sum, RBF, ragged sizes."""
    rng = np.random.default_rng(5)
    out = {}
    cases = {
        "ard_ragged": dict(n=1037, d=3, m=131, kern=lambda d: 2.5 * ARDRBFKernel(np.linspace(0.5, 1.5, d)) + const(0.3) * EyeKernel()),
        "rbf_wide": dict(n=515, d=40, m=64, kern=lambda d: oracle.RBFKernel(3.0)),
        "sum_two": dict(n=300, d=6, m=17, kern=lambda d: 1.5 * ARDRBFKernel(np.full(d, 0.7)) + 0.5 * oracle.RBFKernel(2.0) + const(1) * EyeKernel()),
    }
    for name, c in cases.items():
        X = rng.standard_normal((c["n"], c["d"]))
        y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(c["n"])
        Z = X[rng.permutation(c["n"])[:c["m"]]]
        factory = oracle.get_kernel(lambda c=c: c["kern"](c["d"]), 1e-3)
        theta = factory().get_hyperparameters()
        experts = oracle.get_expert_labels_and_kernels(X, y, factory, 100)
        pred, G, b = oracle.projected_process(experts, Z, factory, theta)
        Xt = rng.standard_normal((50, c["d"]))
        mean, var = pred.predict_many(Xt)
        for k, v in dict(X=X, y=y, Z=Z, Xtest=Xt, G=G, b=b, magic_vector=pred.magic_vector,
                         magic_matrix=pred.magic_matrix, mean=mean, var=var).items():
            out[name + "/" + k] = v
    np.savez_compressed(os.path.join(OUT, "small_cases.npz"), **out)
    print("small cases:", list(cases))


def mnist68(n_rows=1500, m=500, n_e=100):
    """BASELINE config 3 (mnist68 binary GP classification, active=500, RBFKernel(10), tol=1e-3 as
    classification/examples/MNIST.scala:22-32), on the first `n_rows` rows (the full 11 769 x 784 set is 36 MB --
    too large for a fixture).  Features are standardised with the FULL data set's mean / population std
    (commons/util/Scaling.scala); stored as the raw uint8 pixels + the two 784-vectors.  The classification-specific
    step -- the per-expert Laplace mode f (GaussianProcessClassifier.scala:74-129, run once at the initial theta as
    GPCls:60 does after optimisation) -- comes from the oracle; the hot path then sees y := f (GPCls:62-65)."""
    raw = np.loadtxt(os.path.join(REF_DATA, "mnist68.csv"), delimiter=",")
    labels_raw, pix = raw[:, 0], raw[:, 1:]
    n = float(len(pix))
    mean = pix.sum(0) / n
    var = ((pix - mean) ** 2).sum(0) / n
    std = np.sqrt(np.where(var > 0, var, 1.0))
    first = labels_raw[0]                                    # labels201: distinct().collect().zipWithIndex order is
    y01 = (labels_raw != first).astype(np.float64)           # unspecified in Spark; we fix "first seen -> 0"
    P8 = pix[:n_rows + 100].astype(np.uint8)
    X = (P8.astype(np.float64) - mean) / std
    Xtr, ytr, Xte = X[:n_rows], y01[:n_rows], X[n_rows:n_rows + 100]
    from oracle.classification import classification_likelihood_and_gradient
    from oracle import RBFKernel
    factory = oracle.get_kernel(lambda: RBFKernel(10), 1e-3)
    theta = factory().get_hyperparameters()
    experts = oracle.get_expert_labels_and_kernels(Xtr, ytr, factory, n_e)
    groups = oracle.group_for_experts(n_rows, n_e)
    f_all = np.zeros(n_rows)
    negLogZ = 0.0
    fk = []
    for (ye, ke), idx in zip(experts, groups):
        f = np.zeros(len(ye))
        nl, _ = classification_likelihood_and_gradient(ye, f, ke, theta, 1e-3)
        negLogZ += nl
        f_all[idx] = f
        fk.append((f, ke))
    rng = np.random.default_rng(68)
    active_idx = np.sort(rng.permutation(n_rows)[:m])
    Z = Xtr[active_idx]
    pred, G, b = oracle.projected_process(fk, Z, factory, theta)
    fstar, var = pred.predict_many(Xte)
    np.savez_compressed(os.path.join(OUT, "mnist68_case.npz"), pixels=P8, mean=mean, std=std, y01=y01[:n_rows + 100],
                        n_rows=n_rows, f=f_all, active_idx=active_idx, sigma=10.0, sigma2=1e-3, G_diag=np.diag(G).copy(),
                        G_row0=G[0].copy(), b=b, fstar=fstar, var=var, neg_log_z=negLogZ)
    acc = np.mean((fstar > 0) == (y01[n_rows:n_rows + 100] > 0.5))
    print("mnist68: n=%d d=%d m=%d  -logZ=%.4f  hold-out accuracy of sign(f*)=%.2f  |f*|max=%.3f" %
          (n_rows, X.shape[1], m, negLogZ, acc, np.abs(fstar).max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["airfoil", "small", "mnist68"]
    if "airfoil" in which: airfoil()
    if "small" in which: small_synthetic()
    if "mnist68" in which: mnist68()
