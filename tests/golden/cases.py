"""
Case tables shared by tests/golden/make_golden.py (runs the REFERENCE under the numpy shim, build container only) and
tests/test_oracle_golden.py (runs the ORACLE against the stored outputs; no reference needed).
"""
import numpy as np

# ---- direct signature_algs cases: (name, function, M-shape, kwargs) -------------------------------------------------
ALG_CASES = [
    ("fo4_diff", "signature_kern_first_order", (3, 6, 4, 5), dict(num_levels=4, difference=True)),
    ("fo4_nodiff", "signature_kern_first_order", (3, 6, 4, 5), dict(num_levels=4, difference=False)),
    ("fo3_diff", "signature_kern_first_order", (4, 6, 6), dict(num_levels=5, difference=True)),
    ("fo4_lvl1", "signature_kern_first_order", (2, 3, 2, 4), dict(num_levels=1, difference=True)),
    ("ho4_o2", "signature_kern_higher_order", (3, 6, 4, 5), dict(num_levels=4, order=2, difference=True)),
    ("ho4_o3", "signature_kern_higher_order", (3, 6, 4, 5), dict(num_levels=5, order=3, difference=True)),
    ("ho4_o4_nodiff", "signature_kern_higher_order", (2, 5, 3, 5), dict(num_levels=4, order=4, difference=False)),
    ("ho3_o3", "signature_kern_higher_order", (4, 6, 6), dict(num_levels=4, order=3, difference=True)),
    ("tk", "tensor_kern", (10, 5, 6), dict(num_levels=4)),
    ("tvs_fo", "signature_kern_tens_vs_seq_first_order", (10, 4, 3, 7), dict(num_levels=4, difference=True)),
    ("tvs_fo_nodiff", "signature_kern_tens_vs_seq_first_order", (6, 4, 3, 7), dict(num_levels=3, difference=False)),
    ("tvs_ho2", "signature_kern_tens_vs_seq_higher_order", (10, 4, 3, 7), dict(num_levels=4, order=2, difference=True)),
    ("tvs_ho4", "signature_kern_tens_vs_seq_higher_order", (15, 4, 3, 7), dict(num_levels=5, order=4, difference=True)),
]

# ---- SignatureKernel cases ------------------------------------------------------------------------------------------
_ls3 = [0.7, 1.3, 2.1]
_var = lambda M: list(0.5 + 0.25 * np.arange(M + 1))  # noqa: E731

KERNEL_CASES = [
    dict(name="lin_o1", cls="SignatureLinear", kind="linear", L=7, d=3, M=4,
         kw=dict(order=1, normalization=True, lengthscales=_ls3, variances=_var(4)), sigma=1.7),
    dict(name="lin_o1_nonorm", cls="SignatureLinear", kind="linear", L=7, d=3, M=4,
         kw=dict(order=1, normalization=False, lengthscales=_ls3, variances=_var(4)), sigma=0.6),
    dict(name="lin_oM_nonorm", cls="SignatureLinear", kind="linear", L=6, d=3, M=4,
         kw=dict(order=4, normalization=False, lengthscales=None), sigma=1.0),
    dict(name="lin_o2", cls="SignatureLinear", kind="linear", L=7, d=3, M=4,
         kw=dict(order=2, normalization=True, lengthscales=_ls3), sigma=1.0),
    dict(name="rbf_o1", cls="SignatureRBF", kind="rbf", L=7, d=3, M=3,
         kw=dict(order=1, normalization=True, lengthscales=_ls3, variances=_var(3)), sigma=2.0),
    dict(name="rbf_o1_nonorm", cls="SignatureRBF", kind="rbf", L=7, d=3, M=3,
         kw=dict(order=1, normalization=False, lengthscales=_ls3), sigma=1.0),
    dict(name="rbf_o3", cls="SignatureRBF", kind="rbf", L=6, d=2, M=4,
         kw=dict(order=3, normalization=True, lengthscales=[1.5, 0.8]), sigma=1.0),
    dict(name="rbf_nodiff", cls="SignatureRBF", kind="rbf", L=5, d=2, M=3,
         kw=dict(order=1, normalization=True, difference=False, lengthscales=[1.5, 0.8]), sigma=1.0),
    dict(name="rbf_lags", cls="SignatureRBF", kind="rbf", L=7, d=2, M=3,
         kw=dict(order=1, normalization=True, lengthscales=[1.5, 0.8], num_lags=2), sigma=1.0),
    dict(name="lin_lags_nonorm", cls="SignatureLinear", kind="linear", L=6, d=2, M=3,
         kw=dict(order=1, normalization=False, lengthscales=[1.5, 0.8], num_lags=1), sigma=1.0),
    dict(name="cos", cls="SignatureCosine", kind="cosine", L=6, d=3, M=3, kw=dict(lengthscales=_ls3), sigma=1.0),
    dict(name="poly", cls="SignaturePoly", kind="poly", L=6, d=3, M=3,
         kw=dict(lengthscales=_ls3, gamma=0.8, degree=3), sigma=1.0, static=dict(gamma=0.8, degree=3.0)),
    dict(name="mix", cls="SignatureMix", kind="mix", L=6, d=3, M=3, kw=dict(lengthscales=_ls3), sigma=1.0,
         static=dict(mixing=0.5)),
    dict(name="m12", cls="SignatureMatern12", kind="matern12", L=6, d=3, M=3, kw=dict(lengthscales=_ls3), sigma=1.0),
    dict(name="m32", cls="SignatureMatern32", kind="matern32", L=6, d=3, M=3, kw=dict(lengthscales=_ls3), sigma=1.0),
    dict(name="m52", cls="SignatureMatern52", kind="matern52", L=6, d=3, M=3, kw=dict(lengthscales=_ls3), sigma=1.0),
    dict(name="spec_rbf", cls="SignatureSpectral", kind="spectral", L=5, d=2, M=3, kw=dict(family="gauss", Q=3),
         sigma=1.0, spectral=True),
    dict(name="spec_exp", cls="SignatureSpectral", kind="spectral", L=5, d=2, M=3, kw=dict(family="exp", Q=2),
         sigma=1.0, spectral=True),
]

N1, N2, NZ = 5, 4, 6


def kernel_inputs(case, seed=0):
    """Deterministic inputs for one kernel case: sequences X, X2, inducing tensors Z / Zincr, inducing sequences ZS, W."""
    rng = np.random.default_rng([seed, abs(hash(case["name"])) % (2 ** 31)] if False else seed + len(case["name"]) * 7919)
    L, d, M = case["L"], case["d"], case["M"]
    lagmul = (case["kw"].get("num_lags") or 0) + 1
    T = M * (M + 1) // 2
    X = np.cumsum(rng.standard_normal((N1, L, d)), axis=1) / np.sqrt(L)
    X2 = np.cumsum(rng.standard_normal((N2, L, d)), axis=1) / np.sqrt(L)
    Z = 0.6 * rng.standard_normal((T, NZ, d * lagmul))
    Zi = 0.6 * rng.standard_normal((T, NZ, 2, d * lagmul))
    ZS = np.cumsum(rng.standard_normal((NZ, L - 2, d)), axis=1) / np.sqrt(L)
    W = np.eye(NZ)[None] + 0.1 * rng.standard_normal((M, NZ, NZ))
    return dict(X=X.reshape(N1, -1), X2=X2.reshape(N2, -1), Z=Z, Zi=Zi, ZS=ZS.reshape(NZ, -1), W=W)
