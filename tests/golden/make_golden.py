#!/usr/bin/env python
"""
Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference/gpsig) on the numpy-backed
TensorFlow/GPflow stand-ins in tools/refshim.  Build-container only (the GPU box has no /root/reference).

    python tests/golden/make_golden.py

Everything stored is float64.  Randomness of the low-rank path is the shim's deterministic, logged numpy streams
(TF's own streams are unavailable); the golden files carry the draws so the oracle / CUDA path can be fed the same.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
sys.path.insert(0, HERE)

from load_reference import load  # noqa: E402
import cases  # noqa: E402

mods, tf = load()
sa, kr, iv, lr = mods["signature_algs"], mods["kernels"], mods["inducing_variables"], mods["low_rank_calculations"]
T = tf.Tensor


def npy(x):
    if isinstance(x, (tuple, list)):
        return [npy(v) for v in x]
    return np.asarray(tf._a(x), dtype=np.float64)


def gen_algs():
    out = {}
    rng = np.random.default_rng(1234)
    for name, fn, shape, kw in cases.ALG_CASES:
        M = rng.standard_normal(shape)
        out[name + ".M"] = M
        out[name + ".K"] = npy(getattr(sa, fn)(T(M), **kw))
    np.savez_compressed(os.path.join(HERE, "algs.npz"), **out)
    return len(out)


def build_ref_kernel(case):
    L, d, M = case["L"], case["d"], case["M"]
    np.random.seed(4321)  # SignatureSpectral draws alpha/omega/gamma from np.random in __init__ (kernels.py:913-915)
    k = getattr(kr, case["cls"])(L * d, d, M, **case["kw"])
    k.sigma = T(np.float64(case["sigma"]))
    return k


def gen_kernels():
    out = {}
    for case in cases.KERNEL_CASES:
        nm = case["name"]
        k = build_ref_kernel(case)
        inp = cases.kernel_inputs(case)
        for key, v in inp.items():
            out["%s.in.%s" % (nm, key)] = v
        if case.get("spectral"):
            for p in ("alpha", "omega", "gamma"):
                out["%s.param.%s" % (nm, p)] = npy(getattr(k, p))
        X, X2, Z, Zi, ZS, W = (T(inp[q]) for q in ("X", "X2", "Z", "Zi", "ZS", "W"))
        norm = k.normalization
        out[nm + ".K_symm"] = npy(k.K(X))
        out[nm + ".K_symm_lv"] = npy(k.K(X, return_levels=True))
        if case.get("spectral"):
            # the reference's _spectral only handles 2-D inputs (kernels.py:923-925): every path that builds a batched
            # Gram (_K_seq_diag, _K_tens) fails there, so only the symmetric K exists for this kernel
            continue
        out[nm + ".K_rect"] = npy(k.K(X, X2))
        out[nm + ".K_rect_lv"] = npy(k.K(X, X2, return_levels=True))
        out[nm + ".Kdiag"] = npy(k.Kdiag(X))
        out[nm + ".Kdiag_lv"] = npy(k.Kdiag(X, return_levels=True))
        out[nm + ".K_tens"] = npy(k.K_tens(Z))
        out[nm + ".K_tens_lv"] = npy(k.K_tens(Z, return_levels=True))
        out[nm + ".K_itens"] = npy(k.K_tens(Zi, increments=True))
        out[nm + ".K_tvs"] = npy(k.K_tens_vs_seq(Z, X))
        out[nm + ".K_tvs_lv"] = npy(k.K_tens_vs_seq(Z, X, return_levels=True))
        out[nm + ".K_itvs"] = npy(k.K_tens_vs_seq(Zi, X, increments=True))
        for full in (False, True):
            for inc, zz in ((False, Z), (True, Zi)):
                r = npy(k.K_tens_n_seq_covs(zz, X, full_X_cov=full, increments=inc))
                tag = "%s.covs_%s_%s" % (nm, "full" if full else "diag", "inc" if inc else "pts")
                out[tag + ".zz"], out[tag + ".zx"], out[tag + ".xx"] = r
        r = npy(k.K_tens_n_seq_covs(Zi, X, full_X_cov=False, increments=True, return_levels=True))
        out[nm + ".covs_diag_inc_lv.zz"], out[nm + ".covs_diag_inc_lv.zx"], out[nm + ".covs_diag_inc_lv.xx"] = r
        if (case["kw"].get("num_lags") or 0) == 0:
            # inducing sequences (kernels.py:673-761); the full/normalised branch is Q2 (NameError) -> skip it
            r = npy(k.K_seq_n_seq_covs(ZS, X))
            out[nm + ".seqcovs_diag.zz"], out[nm + ".seqcovs_diag.zx"], out[nm + ".seqcovs_diag.xx"] = r
            if not norm:
                r = npy(k.K_seq_n_seq_covs(ZS, X, full_X2_cov=True))
                out[nm + ".seqcovs_full.zz"], out[nm + ".seqcovs_full.zx"], out[nm + ".seqcovs_full.xx"] = r
        # inducing-variable dispatchers (inducing_variables.py:51-137)
        for inc, zz, tag in ((False, inp["Z"], "pts"), (True, inp["Zi"], "inc")):
            for lw in (False, True):
                feat = iv.InducingTensors(zz, case["M"], increments=inc, learn_weights=lw)
                if lw:
                    feat.W = W
                t = "%s.ind_%s_%s" % (nm, tag, "W" if lw else "noW")
                out[t + ".Kuu"] = npy(iv.Kuu(feat, k, jitter=1e-6))
                out[t + ".Kuf"] = npy(iv.Kuf(feat, k, X))
                r = npy(iv.Kuu_Kuf_Kff(feat, k, X, jitter=1e-6, full_f_cov=False))
                out[t + ".zz"], out[t + ".zx"], out[t + ".xx"] = r
        if (case["kw"].get("num_lags") or 0) == 0:
            for lw in (False, True):
                feat = iv.InducingSequences(inp["ZS"], case["M"], learn_weights=lw)
                if lw:
                    feat.W = W
                t = "%s.indseq_%s" % (nm, "W" if lw else "noW")
                out[t + ".Kuu"] = npy(iv.Kuu(feat, k, jitter=1e-6))
                out[t + ".Kuf"] = npy(iv.Kuf(feat, k, X))
                r = npy(iv.Kuu_Kuf_Kff(feat, k, X, jitter=1e-6, full_f_cov=False))
                out[t + ".zz"], out[t + ".zx"], out[t + ".xx"] = r
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    return len(out)


def gen_lowrank():
    """low_rank_calculations.py + the *_lr_feature recursions with seeded (stateless) draws; draws are stored."""
    out = {}
    rng = np.random.default_rng(99)
    A, B = rng.standard_normal((4, 5, 6)), rng.standard_normal((4, 5, 7))
    for sp in ("sqrt", "log"):
        seed = np.array([11, 22], dtype=np.int32)
        del tf.draw_log[:]
        C = npy(lr.lr_hadamard_prod_sparse(T(A), T(B), 8, sp, seed=seed))
        u = dict(tf.draw_log)["sl_uniform"]
        g = dict(tf.draw_log)["sl_normal"]
        D = 6 * 7
        s = D / np.log(D) if sp == "log" else np.sqrt(D)
        out["sparse_%s.A" % sp], out["sparse_%s.B" % sp] = A, B
        out["sparse_%s.R" % sp] = np.where(u <= 1.0 / s, g, 0.0).reshape(D, 8)
        out["sparse_%s.C" % sp] = C
    # subsample ('lin'): tf.random_shuffle of the combination table + Rademacher signs
    del tf.draw_log[:]
    tf.set_shim_seed(5)
    C = npy(lr.lr_hadamard_prod_subsample(T(A), T(B), 9, seed=np.array([3, 4], dtype=np.int32)))
    log = dict(tf.draw_log)
    k1, k2 = 6, 7
    comb = np.stack([np.tile(np.arange(k1), k2), np.repeat(np.arange(k2), k1)], axis=1)[log["perm"]][:9]
    out["subsample.A"], out["subsample.B"], out["subsample.C"] = A, B, C
    out["subsample.select"] = comb.astype(np.float64)
    out["subsample.signs"] = np.where(log["sl_uniform"] <= 0.5, 1.0, -1.0)
    # Nystrom with given landmarks (random diagonal jitter draw logged)
    k = kr.SignatureRBF(6, 2, 3)
    Xp, S = rng.standard_normal((30, 2)), rng.standard_normal((5, 2))
    del tf.draw_log[:]
    F = npy(lr.Nystrom_map(T(Xp), k._base_kern, T(S), None))
    out["nys.X"], out["nys.S"], out["nys.F"] = Xp, S, F
    out["nys.diag_draw"] = dict(tf.draw_log)["uniform"]
    # sequence / tensor low-rank features with seeds (signature_algs.py:162-222)
    U = rng.standard_normal((3, 6, 4))
    seeds = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.int32)
    del tf.draw_log[:]
    Phi = npy(sa.signature_kern_first_order_lr_feature(T(U), 4, 5, "sqrt", T(seeds), difference=True))
    out["lrseq.U"] = U
    draws = [v for kk, v in tf.draw_log]
    for i in range(3):
        u, g = draws[2 * i], draws[2 * i + 1]
        Dm = u.size // 5
        out["lrseq.R%d" % i] = np.where(u <= 1.0 / np.sqrt(Dm), g, 0.0).reshape(Dm, 5)
    for m, P in enumerate(Phi):
        out["lrseq.Phi%d" % m] = P
    Ut = rng.standard_normal((6, 4, 4))
    del tf.draw_log[:]
    Phi = npy(sa.tensor_kern_lr_feature(T(Ut), 3, 5, "sqrt", T(seeds)))
    out["lrtens.U"] = Ut
    draws = [v for kk, v in tf.draw_log]
    for i in range(len(draws) // 2):
        u, g = draws[2 * i], draws[2 * i + 1]
        Dm = u.size // 5
        out["lrtens.R%d" % i] = np.where(u <= 1.0 / np.sqrt(Dm), g, 0.0).reshape(Dm, 5)
    for m, P in enumerate(Phi):
        out["lrtens.Phi%d" % m] = P
    np.savez_compressed(os.path.join(HERE, "lowrank.npz"), **out)
    return len(out)


def walks(n, L, d, seed):
    rng = np.random.default_rng(seed)
    return (np.cumsum(rng.standard_normal((n, L, d)), axis=1) / np.sqrt(L)).reshape(n, L * d)


def gen_configs():
    """BASELINE.json's configurations at sizes the reference (on the NumPy stand-in) finishes in seconds: configs[0] in
    full (SignatureRBF K(X,X) N=32 L=20 d=3 M=3, plus the notebook's order=M linear case), and the tile shapes of
    configs[1] (L=64 d=6 M=4), configs[2-3] (L=128 d=8 M=5; Kuf with incremental inducing tensors) on a few sequences."""
    out = {}

    def put(tag, cls, L, d, M, n, seed, nz=0, **kw):
        k = getattr(kr, cls)(L * d, d, M, **kw)
        X = walks(n, L, d, seed)
        out[tag + ".X"] = X
        out[tag + ".K"] = npy(k.K(T(X)))
        out[tag + ".K_lv"] = npy(k.K(T(X), return_levels=True))
        X2 = walks(max(2, n // 3), L, d, seed + 1)
        out[tag + ".X2"] = X2
        out[tag + ".K_rect"] = npy(k.K(T(X), T(X2)))
        if nz:
            Z = 0.4 * np.random.default_rng(seed + 2).standard_normal((M * (M + 1) // 2, nz, 2, d))
            out[tag + ".Z"] = Z
            r = npy(k.K_tens_n_seq_covs(T(Z), T(X), full_X_cov=False, increments=True))
            out[tag + ".Kzz"], out[tag + ".Kzx"], out[tag + ".Kxx"] = r

    ls3 = [0.9, 1.1, 1.4]
    put("cfg1_rbf", "SignatureRBF", 20, 3, 3, 32, 0, lengthscales=ls3)
    put("cfg1_rbf_nonorm", "SignatureRBF", 20, 3, 3, 32, 0, lengthscales=ls3, normalization=False)
    put("cfg1_lin_orderM", "SignatureLinear", 20, 3, 3, 32, 0, order=3, normalization=False, lengthscales=None)
    put("cfg2_lin_tile", "SignatureLinear", 64, 6, 4, 12, 1, lengthscales=1.0)
    put("cfg4_rbf_tile", "SignatureRBF", 128, 8, 5, 6, 2, nz=8, lengthscales=float(np.sqrt(8.0)))
    put("cfg4_lin_tile", "SignatureLinear", 128, 8, 5, 6, 2, nz=8, lengthscales=1.0)
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **out)
    return len(out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "configs":
        print("configs:", gen_configs(), "arrays")
        print("configs.npz", os.path.getsize(os.path.join(HERE, "configs.npz")), "bytes")
        sys.exit(0)
    print("algs:", gen_algs(), "arrays")
    print("kernels:", gen_kernels(), "arrays")
    print("lowrank:", gen_lowrank(), "arrays")
    print("configs:", gen_configs(), "arrays")
    for f in ("algs.npz", "kernels.npz", "lowrank.npz", "configs.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
