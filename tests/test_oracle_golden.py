"""
Pins oracle/gpsig_oracle.py against outputs of the UNMODIFIED reference sources (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference/gpsig executed on the numpy TF/GPflow stand-in).  CPU only.
"""
import os

import numpy as np
import pytest

import cases
from oracle import gpsig_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALG = np.load(os.path.join(G, "algs.npz"))
KER = np.load(os.path.join(G, "kernels.npz"))
LR = np.load(os.path.join(G, "lowrank.npz"))

TOL = dict(rtol=1e-11, atol=1e-12)  # float64 vs float64, identical op order up to summation inside matmul


@pytest.mark.parametrize("name,fn,shape,kw", cases.ALG_CASES, ids=[c[0] for c in cases.ALG_CASES])
def test_signature_algs_match_reference(name, fn, shape, kw):
    got = getattr(O, fn)(ALG[name + ".M"], **kw)
    np.testing.assert_allclose(got, ALG[name + ".K"], **TOL)


def make_oracle(case):
    kw = dict(case["kw"])
    static = dict(case.get("static", {}))
    for k in ("gamma", "degree", "family", "Q"):
        kw.pop(k, None)
    if case.get("spectral"):
        static = dict(alpha=KER[case["name"] + ".param.alpha"], omega=KER[case["name"] + ".param.omega"],
                      gamma=KER[case["name"] + ".param.gamma"],
                      family="exp" if case["kw"]["family"] == "exp" else "rbf")
        kw["lengthscales"] = None
    return O.SignatureKernelOracle(case["kind"], case["L"] * case["d"], case["d"], case["M"], sigma=case["sigma"],
                                   **kw, **static)


def _inputs(nm):
    return {k: KER["%s.in.%s" % (nm, k)] for k in ("X", "X2", "Z", "Zi", "ZS", "W")}


@pytest.mark.parametrize("case", cases.KERNEL_CASES, ids=[c["name"] for c in cases.KERNEL_CASES])
def test_signature_kernel_matches_reference(case):
    nm = case["name"]
    k = make_oracle(case)
    i = _inputs(nm)
    chk = lambda got, key: np.testing.assert_allclose(got, KER[nm + "." + key], err_msg=key, **TOL)  # noqa: E731
    chk(k.K(i["X"]), "K_symm")
    chk(k.K(i["X"], return_levels=True), "K_symm_lv")
    if case.get("spectral"):
        return
    chk(k.K(i["X"], i["X2"]), "K_rect")
    chk(k.K(i["X"], i["X2"], return_levels=True), "K_rect_lv")
    chk(k.Kdiag(i["X"]), "Kdiag")
    chk(k.Kdiag(i["X"], return_levels=True), "Kdiag_lv")
    chk(k.K_tens(i["Z"]), "K_tens")
    chk(k.K_tens(i["Z"], return_levels=True), "K_tens_lv")
    chk(k.K_tens(i["Zi"], increments=True), "K_itens")
    chk(k.K_tens_vs_seq(i["Z"], i["X"]), "K_tvs")
    chk(k.K_tens_vs_seq(i["Z"], i["X"], return_levels=True), "K_tvs_lv")
    chk(k.K_tens_vs_seq(i["Zi"], i["X"], increments=True), "K_itvs")
    for full in (False, True):
        for inc, zz in ((False, i["Z"]), (True, i["Zi"])):
            r = k.K_tens_n_seq_covs(zz, i["X"], full_X_cov=full, increments=inc)
            tag = "covs_%s_%s" % ("full" if full else "diag", "inc" if inc else "pts")
            for got, part in zip(r, ("zz", "zx", "xx")):
                chk(got, tag + "." + part)
    r = k.K_tens_n_seq_covs(i["Zi"], i["X"], full_X_cov=False, increments=True, return_levels=True)
    for got, part in zip(r, ("zz", "zx", "xx")):
        chk(got, "covs_diag_inc_lv." + part)
    if (case["kw"].get("num_lags") or 0) == 0:
        r = k.K_seq_n_seq_covs(i["ZS"], i["X"])          # literal=True reproduces quirk Q4
        for got, part in zip(r, ("zz", "zx", "xx")):
            chk(got, "seqcovs_diag." + part)
        if not k.normalization:
            r = k.K_seq_n_seq_covs(i["ZS"], i["X"], full_X2_cov=True)
            for got, part in zip(r, ("zz", "zx", "xx")):
                chk(got, "seqcovs_full." + part)


@pytest.mark.parametrize("case", [c for c in cases.KERNEL_CASES if not c.get("spectral")],
                         ids=[c["name"] for c in cases.KERNEL_CASES if not c.get("spectral")])
def test_inducing_dispatch_matches_reference(case):
    nm = case["name"]
    k = make_oracle(case)
    i = _inputs(nm)
    chk = lambda got, key: np.testing.assert_allclose(got, KER[nm + "." + key], err_msg=key, **TOL)  # noqa: E731
    for inc, zz, tag in ((False, i["Z"], "pts"), (True, i["Zi"], "inc")):
        for lw in (False, True):
            W = i["W"] if lw else None
            t = "ind_%s_%s" % (tag, "W" if lw else "noW")
            chk(O.Kuu(k, zz, increments=inc, jitter=1e-6, W=W), t + ".Kuu")
            chk(O.Kuf(k, zz, i["X"], increments=inc, W=W), t + ".Kuf")
            r = O.Kuu_Kuf_Kff(k, zz, i["X"], increments=inc, jitter=1e-6, W=W)
            for got, part in zip(r, ("zz", "zx", "xx")):
                chk(got, t + "." + part)
    if (case["kw"].get("num_lags") or 0) == 0:
        for lw in (False, True):
            W = i["W"] if lw else None
            t = "indseq_%s" % ("W" if lw else "noW")
            chk(O.Kuu(k, i["ZS"], jitter=1e-6, W=W, sequences=True), t + ".Kuu")
            chk(O.Kuf(k, i["ZS"], i["X"], W=W, sequences=True), t + ".Kuf")
            r = O.Kuu_Kuf_Kff(k, i["ZS"], i["X"], jitter=1e-6, W=W, sequences=True)
            for got, part in zip(r, ("zz", "zx", "xx")):
                chk(got, t + "." + part)


def test_lowrank_algebra_matches_reference():
    for sp in ("sqrt", "log"):
        A, B, R = LR["sparse_%s.A" % sp], LR["sparse_%s.B" % sp], LR["sparse_%s.R" % sp]
        got = O.lr_hadamard_prod_sparse(A, B, R, O.sparse_scale(A.shape[-1] * B.shape[-1], sp))
        np.testing.assert_allclose(got, LR["sparse_%s.C" % sp], **TOL)
    got = O.lr_hadamard_prod_subsample(LR["subsample.A"], LR["subsample.B"], LR["subsample.select"].astype(int),
                                       LR["subsample.signs"])
    np.testing.assert_allclose(got, LR["subsample.C"], **TOL)
    rbf = lambda a, b: O.static_kernel("rbf", a, b)  # noqa: E731
    got = O.nystrom_map(LR["nys.X"], rbf, LR["nys.S"], LR["nys.diag_draw"])
    # eigenvector sign/order conventions agree (both numpy eigh) -> direct comparison
    np.testing.assert_allclose(got, LR["nys.F"], rtol=1e-8, atol=1e-10)


def test_lowrank_feature_recursions_match_reference():
    U = LR["lrseq.U"]
    Rs = [LR["lrseq.R%d" % i] for i in range(3)]
    proj = lambda i, A, B: O.lr_hadamard_prod_sparse(A, B, Rs[i], O.sparse_scale(A.shape[-1] * B.shape[-1], "sqrt"))  # noqa: E731
    Phi = O.signature_kern_first_order_lr_feature(U, 4, proj, difference=True, literal=True)
    for m, P in enumerate(Phi):
        np.testing.assert_allclose(P, LR["lrseq.Phi%d" % m], **TOL)
    # quirk Q1: every level >= 2 equals level 1 in the literal restatement
    np.testing.assert_array_equal(Phi[2], Phi[1])
    Ut = LR["lrtens.U"]
    # tensor_kern_lr_feature re-uses seeds[j-1] (signature_algs.py:219): level 2 -> seeds[0]; level 3 -> seeds[0], seeds[1]
    Rt = [LR["lrtens.R%d" % i] for i in range(3)]
    order = iter(Rt)
    proj_t = lambda j, A, B: O.lr_hadamard_prod_sparse(A, B, next(order), O.sparse_scale(A.shape[-1] * B.shape[-1], "sqrt"))  # noqa: E731
    Phi = O.tensor_kern_lr_feature(Ut, 3, proj_t)
    for m, P in enumerate(Phi):
        np.testing.assert_allclose(P, LR["lrtens.Phi%d" % m], **TOL)
