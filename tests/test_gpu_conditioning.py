"""
GPU parity of the RBF paths on data that does NOT sit at unit scale around the first point of the call (round 1 expanded
|x - y|^2 around X[0, 0, :] in fp32 and lost the tolerance 30 lengthscales away).  Reference arithmetic: kernels.py:765-776,
:862-864 in fp64 (the oracle); device: anchored / direct differences in fp32 (warpfused.cu, gram.cu, tens.cu).

Every case runs at the headline tile shape L=128, d=8, M=5, normalised and raw, level by level, tolerance 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import gpsig_oracle as O
from util import assert_close, assert_levels_close, random_walks

pytestmark = pytest.mark.gpu

L, D, M, N = 128, 8, 5, 10


def _cases():
    base = random_walks(N, L, D, 77)
    rng = np.random.default_rng(78)
    out = {}
    for off in (10.0, 30.0, 100.0):
        # the bulk of the data `off` lengthscales away from X[0, 0, :] (which stays where it was: a jump inside sequence 0)
        X = base + off / np.sqrt(D)
        X[0, 0] -= off / np.sqrt(D)
        out["bulk_%g_ls_from_first_point" % off] = (X, 1.0)
        # the same without the jump: everything far from the origin
        out["all_%g_ls_from_origin" % off] = (base + off / np.sqrt(D), 1.0)
    out["per_sequence_offsets"] = (base + 20.0 * rng.standard_normal((N, 1, D)), 1.0)
    out["per_sequence_small_offsets"] = (base + 0.7 * rng.standard_normal((N, 1, D)), 1.0)
    out["amplitude_x20"] = (20.0 * base, 1.0)
    Xt = base.copy()
    Xt[:, :, 0] = np.linspace(0.0, 1.0, L)[None, :]
    ls = np.ones(D)
    ls[0] = 0.01
    out["time_channel_ls_0.01"] = (Xt, ls)
    # small amplitudes: the increments of k ~ 1 - O(amplitude^2) come out of fp32 differences of values next to 1; at 10 % of
    # the lengthscale the tolerance holds with a margin, at 1 % level 1 is at 1.5e-4 .. 3e-4 (accuracy)
    out["amplitude_x0.1"] = (0.1 * base, 1.0)
    return out


CASES = _cases()


def _pair(ls, **kw):
    from gpsig_b200 import kernels
    k = kernels.SignatureRBF(L * D, D, M, lengthscales=ls, **kw)
    ko = O.SignatureKernelOracle("rbf", L * D, D, M, lengthscales=ls, **kw)
    return k, ko


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("norm", [True, False])
def test_K_symm_and_rect_levels(name, norm):
    X, ls = CASES[name]
    X2 = X[::-1][:4].copy()  # rectangular block against copies of some of the sequences, shifted in the last channels (the
    X2[..., 1:] += 0.05      # first may be a time channel with a tiny lengthscale: a shift there leaves nothing but 1e-6 entries)
    Xf, X2f = X.reshape(N, -1), X2.reshape(4, -1)
    k, ko = _pair(ls, normalization=norm)
    assert_levels_close(k.K(Xf, return_levels=True).cpu().numpy(), ko.K(Xf, return_levels=True), msg="%s symm" % name)
    assert_levels_close(k.K(Xf, X2f, return_levels=True).cpu().numpy(), ko.K(Xf, X2f, return_levels=True), msg="%s rect" % name)
    assert_close(k.compute_K_symm(Xf), ko.K(Xf), msg="%s sum" % name)


@pytest.mark.parametrize("name", sorted(CASES))
def test_Kdiag_levels(name):
    X, ls = CASES[name]
    Xf = X.reshape(N, -1)
    k, ko = _pair(ls, normalization=False)
    assert_levels_close(k.Kdiag(Xf, return_levels=True).cpu().numpy(), ko.Kdiag(Xf, return_levels=True), msg="%s diag" % name)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("path", ["fused", "pipeline"])
def test_both_K_paths(name, path):
    """the warp-fused kernel and the two-kernel pipeline (direct differences in the producer) on the same data"""
    from gpsig_b200 import _lib
    X, ls = CASES[name]
    Xf = X.reshape(N, -1)
    k, ko = _pair(ls, normalization=False)
    _lib.set_knob("warpfused", 1 if path == "fused" else 0)
    try:
        got = k.K(Xf, return_levels=True).cpu().numpy()
    finally:
        _lib.set_knob("warpfused", 1)
    assert_levels_close(got, ko.K(Xf, return_levels=True), msg="%s %s" % (name, path))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("increments", [True, False])
def test_Kuf_levels(name, increments):
    """inducing tensors drawn from the data (gpsig/utils.py:25-63: consecutive observations plus noise), so they sit where
    the data sits -- far from the first point of the call in most cases"""
    X, ls = CASES[name]
    Xf = X.reshape(N, -1)
    rng = np.random.default_rng(5)
    T, nz = M * (M + 1) // 2, 7
    seq = rng.integers(0, N, size=(T, nz))
    t = rng.integers(0, L - 1, size=(T, nz))
    scale = np.asarray(ls) * np.ones(D)
    if increments:
        Z = np.stack([X[seq, t], X[seq, t + 1]], axis=2) + 0.1 * scale * rng.standard_normal((T, nz, 2, D))
    else:
        Z = X[seq, t] + 0.1 * scale * rng.standard_normal((T, nz, D))
    for norm in (True, False):
        k, ko = _pair(ls, normalization=norm)
        got = k.K_tens_vs_seq(Z, Xf, increments=increments, return_levels=True).cpu().numpy()
        # a level of Kuf is a sum over time of increments of kernel values (each O(1)); where the data is spread over many
        # lengthscales it cancels to 1e-9 .. 1e-250 in the fp64 reference: absolute floor of a few fp32 roundings of O(1) terms
        assert_levels_close(got, ko.K_tens_vs_seq(Z, Xf, increments=increments, return_levels=True),
                            msg="%s Kuf inc=%s norm=%s" % (name, increments, norm), atol=2e-6)


def test_Kuf_long_tensor_increments_take_the_direct_form():
    """inducing tensors whose own increment is many lengthscales long (free parameters can end up there)"""
    X = random_walks(N, L, D, 3)
    Xf = X.reshape(N, -1)
    rng = np.random.default_rng(6)
    T, nz = M * (M + 1) // 2, 5
    Z = X[rng.integers(0, N, size=(T, nz)), rng.integers(0, L, size=(T, nz))][:, :, None, :] + \
        np.stack([np.zeros((T, nz, D)), 6.0 * rng.standard_normal((T, nz, D))], axis=2)
    k, ko = _pair(1.0, normalization=False)
    got = k.K_tens_vs_seq(Z, Xf, increments=True, return_levels=True).cpu().numpy()
    assert_levels_close(got, ko.K_tens_vs_seq(Z, Xf, increments=True, return_levels=True), msg="long dz")


def test_sharded_rows_and_columns_are_bit_equal_to_the_single_call():
    """no call-level centre any more: row blocks of K(X, X2) and column blocks of Kuf do not depend on what else is in the
    call (world size 1 here: the shards are computed one after the other on this GPU)"""
    X, ls = CASES["per_sequence_small_offsets"]
    Xf = X.reshape(N, -1)
    k, _ = _pair(ls)
    full = k.K(Xf, Xf[:6])
    parts = torch.cat([k.K(Xf[:3], Xf[:6]), k.K(Xf[3:], Xf[:6])], dim=0)
    assert torch.equal(full, parts)
    rng = np.random.default_rng(9)
    Z = X[0, 5][None, None, None, :] + 0.3 * rng.standard_normal((M * (M + 1) // 2, 6, 2, D))
    fullz = k.K_tens_vs_seq(Z, Xf, increments=True)
    partz = torch.cat([k.K_tens_vs_seq(Z, Xf[:4], increments=True), k.K_tens_vs_seq(Z, Xf[4:], increments=True)], dim=1)
    assert torch.equal(fullz, partz)
