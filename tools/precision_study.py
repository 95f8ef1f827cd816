"""Design study (CPU, numpy): how much arithmetic precision do the sufficient statistics G = K_mn K_nm,
b = K_mn y need for the posterior mean / variance to match the fp64 oracle to 1e-5 relative?

Emulates operand roundings a tensor-core Gram would apply (fp16 / bf16 / tf32 / fp16 hi+lo split / fp32
elements) with exact (fp64) accumulation, and fp32 chunk accumulation with fp64 flushes.  Not product
code; results are quoted in DESIGN.md."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import oracle
from oracle import ARDRBFKernel, EyeKernel, const

def synth(N, d, seed=13):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d), dtype=np.float32)
    y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(N)
    return X, y

def round_to(K, kind):
    if kind == "f64": return K
    if kind == "f32": return K.astype(np.float32).astype(np.float64)
    if kind == "f16": return K.astype(np.float16).astype(np.float64)
    if kind == "bf16": return torch.from_numpy(K).to(torch.bfloat16).to(torch.float64).numpy()
    if kind == "tf32":
        b = K.astype(np.float32).view(np.uint32).astype(np.uint64)
        b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)     # round-to-nearest (ties away) to 10 bits
        return b.view(np.float32).astype(np.float64)
    raise ValueError(kind)

def variants(K64):
    out = {}
    out["f64"] = (K64, None)
    K32 = K64.astype(np.float32).astype(np.float64)
    out["f32"] = (K32, None)
    out["f16"] = (round_to(K64, "f16"), None)
    out["bf16"] = (round_to(K64, "bf16"), None)
    out["tf32"] = (round_to(K64, "tf32"), None)
    hi = round_to(K32, "f16"); lo = round_to((K32 - hi) * 2048.0, "f16") / 2048.0
    out["f16x2"] = (hi, lo)            # G ~ hi'hi + hi'lo + lo'hi
    hi = round_to(K32, "bf16"); lo = round_to(K32 - hi, "bf16")
    out["bf16x2"] = (hi, lo)
    hi = round_to(K32, "tf32"); lo = round_to(K32 - hi, "tf32")
    out["tf32x2"] = (hi, lo)
    return out

def main(N=200000, d=16, m=1000, n_test=1000, chunk=20000, q32=True):
    X, y = synth(N + n_test, d)
    Xt, X, y = X[N:], X[:N], y[:N]
    beta = np.full(d, np.sqrt(18.0 / d))
    rng = np.random.default_rng(7)
    Z = X[rng.permutation(N)[:m]].astype(np.float64)
    user = lambda: 1 * ARDRBFKernel(d) + const(1) * EyeKernel()
    factory = oracle.get_kernel(user, 1e-4)
    theta = np.concatenate([[1.0], beta])
    names = ["f64", "f32", "f16", "bf16", "tf32", "f16x2", "bf16x2", "tf32x2", "f32q", "f16+acc32", "f16x2+acc32"]
    G = {n: np.zeros((m, m)) for n in names}
    b = {n: np.zeros(m) for n in names}
    Zb = Z * beta
    zz = (Zb * Zb).sum(1)
    t0 = time.time()
    for s in range(0, N, chunk):
        Xc = X[s:s + chunk].astype(np.float64); yc = y[s:s + chunk]
        Xb = Xc * beta
        q = (Xb * Xb).sum(1)[:, None] + zz[None, :] - 2.0 * Xb @ Zb.T
        q = np.maximum(q, 0.0)
        K64 = np.exp(-q)                                   # n x m
        v = variants(K64)
        for n, (hi, lo) in v.items():
            g = hi.T @ hi
            if lo is not None:
                c = hi.T @ lo
                g = g + c + c.T
            G[n] += g
            b[n] += (hi if lo is None else hi + lo).T @ yc
        # q computed in fp32 (direct-form distance in fp32, exp in fp32), exact Gram
        X32 = X[s:s + chunk]; Z32 = Z.astype(np.float32); b32 = beta.astype(np.float32)
        q32v = np.zeros((len(X32), m), dtype=np.float32)
        for k in range(d):
            df = (X32[:, k][:, None] - Z32[:, k][None, :]) * b32[k]
            q32v += df * df
        Kq = np.exp(-q32v).astype(np.float64)
        G["f32q"] += Kq.T @ Kq; b["f32q"] += Kq.T @ yc
        # fp32 accumulation per 2048-point sub-chunk, flushed to fp64
        for nm, src in (("f16+acc32", v["f16"][0]), ("f16x2+acc32", None)):
            for ss in range(0, len(yc), 2048):
                if src is not None:
                    h = torch.from_numpy(src[ss:ss + 2048].astype(np.float32))
                    g = (h.T @ h).numpy().astype(np.float64)
                    bb = (h.T @ torch.from_numpy(yc[ss:ss + 2048].astype(np.float32))).numpy().astype(np.float64)
                else:
                    h = torch.from_numpy(v["f16x2"][0][ss:ss + 2048].astype(np.float32))
                    l = torch.from_numpy(v["f16x2"][1][ss:ss + 2048].astype(np.float32))
                    c = (h.T @ l)
                    g = ((h.T @ h) + c + c.T).numpy().astype(np.float64)
                    bb = ((h + l).T @ torch.from_numpy(yc[ss:ss + 2048].astype(np.float32))).numpy().astype(np.float64)
                G[nm] += g; b[nm] += bb
    print("stats done in %.1fs" % (time.time() - t0), flush=True)
    kern = factory().set_hyperparameters(theta).set_training_vectors(Z)
    ref = None
    A0 = kern.white_noise_var * kern.training_kernel() + G["f64"]
    ev = np.linalg.eigvalsh(A0)
    print("N=%d d=%d m=%d  cond(A)=%.3e  lam_min=%.4f lam_max=%.3e  max|G|=%.3e" % (N, d, m, ev[-1] / ev[0], ev[0], ev[-1], np.abs(G["f64"]).max()))
    print("%-12s %10s %10s %10s %10s %10s %10s" % ("variant", "dG/maxG", "db/maxb", "mv_rel", "mean_max", "mean_rms", "var_max"))
    for n in names:
        mv, mm = oracle.get_magic_vector(kern, G[n], b[n])
        pred = oracle.GaussianProjectedProcessRawPredictor(mv, mm, kern)
        mean, var = pred.predict_many(Xt.astype(np.float64))
        if ref is None:
            ref = (mv, mean, var)
        dG = np.abs(G[n] - G["f64"]).max() / np.abs(G["f64"]).max()
        db = np.abs(b[n] - b["f64"]).max() / np.abs(b["f64"]).max()
        mvr = np.abs(mv - ref[0]).max() / np.abs(ref[0]).max()
        me = np.abs(mean - ref[1]).max() / np.abs(ref[1]).max()
        mr = np.sqrt(np.mean((mean - ref[1]) ** 2)) / np.sqrt(np.mean(ref[1] ** 2))
        ve = (np.abs(var - ref[2]) / np.abs(ref[2])).max()
        print("%-12s %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % (n, dG, db, mvr, me, mr, ve), flush=True)

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
