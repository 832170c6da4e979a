"""
BASELINE.json's configurations against the REFERENCE's own outputs (tests/golden/configs.npz, produced by
tests/golden/make_golden.py from the unmodified /root/reference sources): configs[0] in full (SignatureRBF K(X,X) N=32
L=20 d=3 M=3 and the notebook's order=M linear kernel), and the tile shapes of configs[1] (L=64 d=6 M=4) and
configs[2-3] (L=128 d=8 M=5, Kuf / Kuu_Kuf_Kff with incremental inducing tensors) on a handful of sequences.
The oracle is held to 1e-10 (CPU), the device path to the north star's 1e-4 (GPU; measured ~1e-7).
"""
import numpy as np
import pytest

from oracle import gpsig_oracle as O
from util import GOLDEN, assert_close, assert_levels_close

CFG = np.load(GOLDEN + "/configs.npz")
CASES = [
    ("cfg1_rbf", "rbf", "SignatureRBF", 20, 3, 3, dict(lengthscales=[0.9, 1.1, 1.4])),
    ("cfg1_rbf_nonorm", "rbf", "SignatureRBF", 20, 3, 3, dict(lengthscales=[0.9, 1.1, 1.4], normalization=False)),
    ("cfg1_lin_orderM", "linear", "SignatureLinear", 20, 3, 3, dict(order=3, normalization=False, lengthscales=None)),
    ("cfg2_lin_tile", "linear", "SignatureLinear", 64, 6, 4, dict(lengthscales=1.0)),
    ("cfg4_rbf_tile", "rbf", "SignatureRBF", 128, 8, 5, dict(lengthscales=float(np.sqrt(8.0)))),
    ("cfg4_lin_tile", "linear", "SignatureLinear", 128, 8, 5, dict(lengthscales=1.0)),
]


@pytest.mark.parametrize("tag,kind,cls,L,d,M,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_on_baseline_configs(tag, kind, cls, L, d, M, kw):
    okw = dict(kw)
    if okw.get("lengthscales") is not None:
        okw["lengthscales"] = np.asarray(okw["lengthscales"], dtype=np.float64) * np.ones(d)
    ko = O.SignatureKernelOracle(kind, L * d, d, M, **okw)
    X, X2 = CFG[tag + ".X"], CFG[tag + ".X2"]
    np.testing.assert_allclose(ko.K(X), CFG[tag + ".K"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ko.K(X, return_levels=True), CFG[tag + ".K_lv"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ko.K(X, X2), CFG[tag + ".K_rect"], rtol=1e-9, atol=1e-11)
    if tag + ".Z" in CFG:
        r = ko.K_tens_n_seq_covs(CFG[tag + ".Z"], X, increments=True)
        for got, part in zip(r, ("Kzz", "Kzx", "Kxx")):
            np.testing.assert_allclose(got, CFG[tag + "." + part], rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kind,cls,L,d,M,kw", CASES, ids=[c[0] for c in CASES])
def test_device_matches_reference_on_baseline_configs(tag, kind, cls, L, d, M, kw):
    from gpsig_b200 import kernels
    k = getattr(kernels, cls)(L * d, d, M, **kw)
    X, X2 = CFG[tag + ".X"], CFG[tag + ".X2"]
    assert_close(k.compute_K_symm(X), CFG[tag + ".K"], msg=tag + ".K")
    assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), CFG[tag + ".K_lv"], msg=tag + ".K_lv")
    assert_close(k.compute_K(X, X2), CFG[tag + ".K_rect"], msg=tag + ".K_rect")
    if tag + ".Z" in CFG:
        r = k.K_tens_n_seq_covs(CFG[tag + ".Z"], X, increments=True)
        for got, part in zip(r, ("Kzz", "Kzx", "Kxx")):
            assert_close(got.cpu().numpy(), CFG[tag + "." + part], msg=tag + "." + part)
