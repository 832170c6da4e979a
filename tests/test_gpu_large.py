"""
GPU parity at sizes past the golden fixtures: BASELINE.json's configs at reduced N against the oracle, the headline
tile shape (L=128, d=8, M=5), chunked workspaces, and size-independent properties at full N.
"""
import numpy as np
import pytest
import torch

from oracle import gpsig_oracle as O
from util import assert_close, assert_levels_close, random_walks

pytestmark = pytest.mark.gpu


def _pair(kind, L, d, M, **kw):
    from gpsig_b200 import kernels
    cls = dict(linear=kernels.SignatureLinear, rbf=kernels.SignatureRBF)[kind]
    k = cls(L * d, d, M, **kw)
    ko = O.SignatureKernelOracle(kind, L * d, d, M, **kw)
    return k, ko


def test_config1_rbf_n32_l20_d3_m3():
    """BASELINE.json configs[0]: SignatureRBF K(X,X) N=32 L=20 d=3 M=3."""
    X = random_walks(32, 20, 3, 0).reshape(32, -1)
    ls = np.array([0.9, 1.1, 1.4])
    for norm in (True, False):
        k, ko = _pair("rbf", 20, 3, 3, lengthscales=ls, normalization=norm)
        assert_close(k.compute_K_symm(X), ko.K(X), msg="cfg1 norm=%s" % norm)
        assert_levels_close(k.compute_K_levels(X, X[:7]), ko.K(X, X[:7], return_levels=True), msg="cfg1 rect")
    k, ko = _pair("linear", 20, 3, 3, order=3, normalization=False, lengthscales=None)
    assert_close(k.compute_K_symm(X), ko.K(X), msg="cfg1 order=M")


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_config2_shape_subsampled(kind):
    """configs[1] tile shape (L=64, d=6, M=4) on 96 sequences (the oracle needs N^2 L^2 doubles)."""
    X = random_walks(96, 64, 6, 1).reshape(96, -1)
    k, ko = _pair(kind, 64, 6, 4)
    assert_close(k.compute_K_symm(X), ko.K(X, row_block=16), msg="cfg2")


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_headline_tile_shape_subsampled(kind):
    """configs[3] tile shape (L=128, d=8, M=5) on 48 sequences, symmetric and rectangular, levels and sum."""
    X = random_walks(48, 128, 8, 2).reshape(48, -1)
    Y = random_walks(21, 128, 8, 3).reshape(21, -1)
    ls = np.sqrt(8.0) * np.ones(8) if kind == "rbf" else np.ones(8)
    k, ko = _pair(kind, 128, 8, 5, lengthscales=ls)
    assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), ko.K(X, return_levels=True, row_block=8), msg="symm")
    assert_close(k.compute_K(X, Y), ko.K(X, Y, row_block=8), msg="rect")
    k2, ko2 = _pair(kind, 128, 8, 5, lengthscales=ls, normalization=False)
    assert_levels_close(k2.K(X, return_levels=True).cpu().numpy(), ko2.K(X, return_levels=True, row_block=8), msg="raw")


def test_chunked_workspace_gives_identical_result():
    """A tiny workspace budget forces many row-block chunks; the result must not change by a single bit."""
    from gpsig_b200 import settings
    X = random_walks(70, 64, 4, 4).reshape(70, -1)
    k, _ = _pair("rbf", 64, 4, 4)
    ref = k.K(X).clone()
    old = settings.workspace_budget_bytes
    try:
        settings.workspace_budget_bytes = 3 << 20
        k._ws = None
        got = k.K(X)
        got_rect = k.K(X, X[:33])
    finally:
        settings.workspace_budget_bytes = old
        k._ws = None
    assert torch.equal(ref, got)
    assert_close(got_rect.cpu().numpy(), ref[:, :33].cpu().numpy(), tol=1e-5)


def test_kuf_shape_subsampled_vs_oracle():
    """configs[2] shape (L=128, d=8, M=5) with 32 inducing tensors and 64 sequences, increments on and off."""
    rng = np.random.default_rng(5)
    X = random_walks(64, 128, 8, 5).reshape(64, -1)
    T = 15
    for inc in (False, True):
        Z = 0.5 * rng.standard_normal((T, 32, 2, 8) if inc else (T, 32, 8))
        for kind in ("linear", "rbf"):
            k, ko = _pair(kind, 128, 8, 5, lengthscales=2.0 * np.ones(8))
            got = k.K_tens_vs_seq(Z, X, increments=inc, return_levels=True).cpu().numpy()
            assert_levels_close(got, ko.K_tens_vs_seq(Z, X, increments=inc, return_levels=True), msg="Kuf %s" % kind)
            Kzz, Kzx, Kxx = k.K_tens_n_seq_covs(Z, X, increments=inc)
            rzz, rzx, rxx = ko.K_tens_n_seq_covs(Z, X, increments=inc)
            assert_close(Kzz.cpu().numpy(), rzz, msg="Kzz")
            assert_close(Kzx.cpu().numpy(), rzx, msg="Kzx")
            assert_close(Kxx.cpu().numpy(), rxx, msg="Kxx")


def test_full_size_properties_n1024():
    """configs[1] at full size (N=1024, L=64, d=6, M=4): symmetry, unit diagonal, positive semi-definiteness (via
    Cholesky of K + 1e-4 I), and a 32x32 corner against the oracle."""
    X = random_walks(1024, 64, 6, 6).reshape(1024, -1)
    k, ko = _pair("linear", 64, 6, 4)
    K = k.K(X)
    assert torch.equal(K, K.T)
    assert torch.allclose(torch.diagonal(K), torch.full((1024,), 5.0, device=K.device), atol=1e-4)
    torch.linalg.cholesky(K.double() + 1e-4 * torch.eye(1024, device=K.device, dtype=torch.float64))
    assert_close(K[:32, :32].cpu().numpy(), ko.K(X[:32]), msg="corner")
    # rectangular call on the same data agrees with the symmetric one off the diagonal
    Kr = k.K(X[:64], X[64:256])
    assert_close(Kr.cpu().numpy(), K[:64, 64:256].cpu().numpy(), tol=1e-5)


def test_svgp_elbo_vs_oracle():
    """gpsig/models.py:39-73 consumer: ELBO and predictive moments on a small problem (fp32 device vs fp64 oracle)."""
    from gpsig_b200 import kernels, inducing_variables as iv, models
    rng = np.random.default_rng(7)
    n, L, d, M, nz = 60, 32, 3, 3, 10
    X = random_walks(n, L, d, 7).reshape(n, -1)
    Y = (rng.standard_normal((n, 1)) > 0).astype(np.float64)
    Z = 0.4 * rng.standard_normal((6, nz, 2, d))
    q_mu = 0.3 * rng.standard_normal((nz, 1))
    q_sqrt = np.tril(0.2 * rng.standard_normal((1, nz, nz))) + np.eye(nz)[None]
    k = kernels.SignatureRBF(L * d, d, M, lengthscales=1.5)
    ko = O.SignatureKernelOracle("rbf", L * d, d, M, lengthscales=1.5)
    feat = iv.InducingTensors(Z, M, increments=True)
    for lik, name in ((models.Gaussian(0.5), "gaussian"), (models.Bernoulli(), "bernoulli")):
        m = models.SVGP(X, Y, k, lik, feat, num_latent=1, q_mu=q_mu, q_sqrt=q_sqrt)
        elbo = m.compute_log_likelihood()
        ref, fm, fv = O.svgp_elbo(ko, Z, X, Y, q_mu, q_sqrt, likelihood=name, lik_variance=0.5, increments=True)
        assert abs(elbo - ref) / abs(ref) < 2e-4, (elbo, ref)
        mu, var = m.predict_f(X)
        assert_close(mu.cpu().numpy(), fm, tol=2e-4, msg="fmean")
        assert_close(var.cpu().numpy(), fv, tol=2e-4, msg="fvar")


@pytest.mark.parametrize("M", [6, 7, 8])
@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_many_levels_on_the_stream_path(kind, M):
    """num_levels 6..8 use the 8-warp instantiation of the stream recursion (register budget); L=40 -> LP=4, G=8."""
    X = random_walks(37, 40, 3, 20 + M).reshape(37, -1)
    Y = random_walks(10, 40, 3, 40 + M).reshape(10, -1)
    k, ko = _pair(kind, 40, 3, M, lengthscales=1.7)
    assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), ko.K(X, return_levels=True), msg="symm M=%d" % M)
    assert_levels_close(k.K(X, Y, return_levels=True).cpu().numpy(), ko.K(X, Y, return_levels=True), msg="rect M=%d" % M)


@pytest.mark.parametrize("L,d,M", [(100, 10, 6), (45, 3, 4), (17, 5, 2), (300, 2, 3), (2, 4, 3)])
@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_sequence_shapes_off_the_headline_tile(kind, L, d, M):
    """config-5 tile (L=100, d=10, M=6), LIBRAS-like odd lengths, long sequences (LP=32), two-point sequences."""
    n = 21
    X = random_walks(n, L, d, L + d).reshape(n, -1)
    k, ko = _pair(kind, L, d, M, lengthscales=0.5 * np.sqrt(d) + 0.5)
    assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), ko.K(X, return_levels=True, row_block=7), msg="symm")
    k2, ko2 = _pair(kind, L, d, M, lengthscales=0.5 * np.sqrt(d) + 0.5, normalization=False)
    assert_close(k2.Kdiag(X).cpu().numpy(), ko2.Kdiag(X), msg="Kdiag")


@pytest.mark.parametrize("kind", ["linear", "rbf"])
@pytest.mark.parametrize("L,d,M,diff", [(100, 10, 6, True), (33, 3, 4, True), (20, 5, 3, False), (64, 16, 2, True),
                                        (12, 2, 7, True)])
def test_kuf_fast_and_generic_paths_vs_oracle(kind, L, d, M, diff):
    """fused tensor-vs-sequence kernel: fast instantiations (M <= 6, d <= 16) and the generic fallback (M = 7)."""
    rng = np.random.default_rng(L * d + M)
    n, nz, T = 45, 11, M * (M + 1) // 2
    X = random_walks(n, L, d, 9).reshape(n, -1)
    for inc in (False, True):
        Z = 0.5 * rng.standard_normal((T, nz, 2, d) if inc else (T, nz, d))
        k, ko = _pair(kind, L, d, M, lengthscales=1.3 * np.ones(d), difference=diff, normalization=False)
        got = k.K_tens_vs_seq(Z, X, increments=inc, return_levels=True).cpu().numpy()
        assert_levels_close(got, ko.K_tens_vs_seq(Z, X, increments=inc, return_levels=True), msg="Kuf %s inc=%s" % (kind, inc))


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_fused_kernel_against_the_pipeline(kind):
    """The warp-fused Gram + recursion kernel (warpfused.cu, default) against the two-kernel path (increment-Gram producer
    -> stream recursion).  Linear: same arithmetic per entry, so the same bits.  RBF: the fused kernel evaluates the
    squared distance in the anchored form, the producer directly from the differences -- equal to fp32 rounding."""
    import ctypes
    from gpsig_b200 import _lib
    X = random_walks(75, 64, 5, 31).reshape(75, -1)
    Y = random_walks(22, 64, 5, 32).reshape(22, -1)
    k, _ = _pair(kind, 64, 5, 4, lengthscales=1.4)
    lib = _lib.load()
    _lib.set_knob("warpfused", 0)
    try:
        ref_s, ref_r = k.K(X, return_levels=True).clone(), k.K(X, Y, return_levels=True).clone()
    finally:
        _lib.set_knob("warpfused", 1)
    lib.gpsig_profile_reset(); lib.gpsig_profile_enable(1)
    got_s, got_r = k.K(X, return_levels=True), k.K(X, Y, return_levels=True)
    torch.cuda.synchronize()
    lib.gpsig_profile_enable(0)
    ms, n, un = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.gpsig_profile_read(6, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(un))
    lib.gpsig_profile_reset()
    assert n.value >= 2, "the fused kernel did not run"
    if kind == "linear":
        assert torch.equal(got_s, ref_s) and torch.equal(got_r, ref_r)
    else:
        for got, ref in ((got_s, ref_s), (got_r, ref_r)):
            for m in range(got.shape[0]):
                assert float((got[m] - ref[m]).abs().max()) <= 5e-6 * float(ref[m].abs().max()), m


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_edge_shapes(kind):
    """one sequence; two-point sequences; one level; ragged pair groups; sequences beyond the 512-column fast path."""
    for n, L, d, M in ((1, 16, 2, 3), (3, 2, 2, 2), (5, 9, 1, 1), (7, 33, 3, 2), (3, 600, 2, 2)):
        X = random_walks(n, L, d, 100 + L).reshape(n, -1)
        k, ko = _pair(kind, L, d, M, lengthscales=1.1)
        assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), ko.K(X, return_levels=True), msg="symm %s" % ((n, L, d, M),))
        Y = random_walks(2, L, d, 200 + L).reshape(2, -1)
        assert_close(k.compute_K(X, Y), ko.K(X, Y), msg="rect %s" % ((n, L, d, M),))
    # a single time step: no increments at all -> every level >= 1 vanishes, K = variances[0] * sigma (normalised: 0/0 guarded
    # by the jitter exactly as in the reference)
    X1 = random_walks(4, 1, 3, 5).reshape(4, -1)
    k, ko = _pair(kind, 1, 3, 3, normalization=False)
    assert_close(k.compute_K_symm(X1), ko.K(X1), msg="L=1")


def test_unsupported_configurations_raise_instead_of_falling_back():
    from gpsig_b200 import kernels, _lib
    with pytest.raises(ValueError):
        kernels.SignatureRBF(10, 3, 2)                           # input_dim not a multiple of num_features (kernels.py:98-101)
    with pytest.raises(NotImplementedError):
        kernels.SignatureSpectral(8, 2, 2, family="mixed")


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_wide_state_space_falls_back_to_the_materialised_gram(kind):
    """d = 24 (> 16): Gram blocks through gpsig_gram + the operator-level recursion; same results, slower path."""
    n, L, d, M = 9, 32, 24, 3
    X = random_walks(n, L, d, 77).reshape(n, -1)
    Y = random_walks(4, L, d, 78).reshape(4, -1)
    k, ko = _pair(kind, L, d, M, lengthscales=float(np.sqrt(d)))
    assert_levels_close(k.K(X, return_levels=True).cpu().numpy(), ko.K(X, return_levels=True), msg="symm")
    assert_close(k.compute_K(X, Y), ko.K(X, Y), msg="rect")
    rng = np.random.default_rng(3)
    Z = 0.3 * rng.standard_normal((6, 5, 2, d))
    r, ro = k.K_tens_n_seq_covs(Z, X, increments=True), ko.K_tens_n_seq_covs(Z, X, increments=True)
    for a, b, nm in zip(r, ro, ("zz", "zx", "xx")):
        assert_close(a.cpu().numpy(), b, msg=nm)


def test_gram_with_more_rows_than_a_grid_dimension():
    """gpsig_gram put the rows on gridDim.y (cap 65535 blocks of 8): the low-rank mode's Nystrom map calls it
    with N L rows.  700k rows against 50 landmarks."""
    from gpsig_b200 import kernels
    rng = np.random.default_rng(0)
    A = rng.standard_normal((700_000, 3))
    B = rng.standard_normal((50, 3))
    k = kernels.SignatureRBF(3, 3, 2, lengthscales=None)
    G = k._base_gram(torch.as_tensor(A, dtype=torch.float32, device="cuda"), torch.as_tensor(B, dtype=torch.float32, device="cuda"))
    rows = np.array([0, 1, 65535 * 8 - 1, 65535 * 8, 65535 * 8 + 7, 699_999])
    ref = np.exp(-0.5 * ((A[rows][:, None, :] - B[None, :, :]) ** 2).sum(-1))
    assert_close(G[torch.as_tensor(rows, device="cuda")].cpu().numpy(), ref, tol=1e-5, msg="gram rows")


def test_diag_pass_of_many_short_sequences_on_a_fresh_kernel():
    """the diagonal pass keeps the prepared points of ALL sequences in its workspace; a fresh kernel object (no
    cached workspace from an earlier K call) with 20000 sequences of 10 points."""
    n, L, d, M = 20000, 10, 4, 3
    X = random_walks(n, L, d, 5).reshape(n, -1)
    for kind in ("rbf", "linear"):
        k, ko = _pair(kind, L, d, M, normalization=False)
        got = k.Kdiag(X, return_levels=True).cpu().numpy()
        sel = np.array([0, 1, 7777, n - 1])
        assert_levels_close(got[:, sel], ko.Kdiag(X[sel], return_levels=True), msg="diag %s" % kind)
        k2, _ = _pair(kind, L, d, M)
        Z = 0.5 * np.random.default_rng(1).standard_normal((M * (M + 1) // 2, 3, 2, d))
        assert k2.K_tens_vs_seq(Z, X, increments=True).shape == (3, n)


def test_tcgen05_kuf_many_items_per_cta_and_a_single_sequence_chunk():
    """The tcgen05 Kuf kernel is persistent: with 8 row tiles x 21 sequence chunks (168 items > 148 SMs) some CTAs take two
    items, and the last chunk holds ONE sequence, so the second warp set's producer has no tiles in it -- it must still
    stay in step with the producer that loads the A tiles (it used to skip that wait and could run an item ahead)."""
    from gpsig_b200 import _lib
    rng = np.random.default_rng(11)
    L, d, M, nz, n = 64, 8, 5, 64, 161
    X = random_walks(n, L, d, 9).reshape(n, -1)
    T = M * (M + 1) // 2
    Xr = X.reshape(n, L, d)
    seq, t = rng.integers(0, n, size=(T, nz)), rng.integers(0, L - 1, size=(T, nz))
    Z = np.stack([Xr[seq, t], Xr[seq, t + 1]], axis=2) + 0.3 * rng.standard_normal((T, nz, 2, d))
    k, ko = _pair("rbf", L, d, M, lengthscales=float(np.sqrt(d)))
    got = k.K_tens_vs_seq(Z, X, increments=True, return_levels=True).cpu().numpy()
    _lib.set_knob("tens_tc", 0)
    try:
        cuda_core = k.K_tens_vs_seq(Z, X, increments=True, return_levels=True).cpu().numpy()
    finally:
        _lib.set_knob("tens_tc", 1)
    assert_levels_close(got, cuda_core, msg="tcgen05 vs CUDA-core Kuf")
    cols = np.r_[0:4, 80:84, 157:161]      # first chunk, a chunk some CTA takes as its second item, the one-sequence chunk
    ref = ko.K_tens_vs_seq(Z[:, :6], X[cols], increments=True, return_levels=True)
    assert_levels_close(got[:, :6][:, :, cols], ref, msg="tcgen05 Kuf vs oracle")


@pytest.mark.parametrize("L,d,M,nz,n", [(2, 3, 2, 1, 1), (3, 8, 5, 2, 3), (17, 9, 4, 5, 7), (65, 10, 6, 3, 2), (130, 12, 3, 9, 5)])
def test_tcgen05_kuf_edge_shapes(L, d, M, nz, n):
    """Shortest sequences (one or two time steps of increments), one tensor / one sequence, a single block in the last
    tile (L = 65, 130), and the three operand layouts (d <= 8; d = 9, 10; d = 12) of the tcgen05 Kuf kernel."""
    rng = np.random.default_rng(L * 31 + d)
    X = random_walks(n, L, d, 3).reshape(n, -1)
    T = M * (M + 1) // 2
    Xr = X.reshape(n, L, d)
    seq, t = rng.integers(0, n, size=(T, nz)), rng.integers(0, L - 1, size=(T, nz))
    Z = np.stack([Xr[seq, t], Xr[seq, t + 1]], axis=2) + 0.2 * rng.standard_normal((T, nz, 2, d))
    k, ko = _pair("rbf", L, d, M, lengthscales=float(np.sqrt(d)), normalization=False)
    got = k.K_tens_vs_seq(Z, X, increments=True, return_levels=True).cpu().numpy()
    assert_levels_close(got, ko.K_tens_vs_seq(Z, X, increments=True, return_levels=True), msg="Kuf edge L=%d d=%d" % (L, d))
