"""
GPU: reverse mode of the covariance path.

  * the two CUDA VJP kernels (csrc/vjp.cu) against the oracle's reverse-mode restatement and against central finite
    differences of the fp64 oracle;
  * gradients of the kernel-level covariances (Kzz, Kzx, Kxx of K_tens_n_seq_covs; K; Kdiag) with respect to lengthscales,
    variances, inducing tensors and inputs against finite differences of the oracle -- tolerance 1e-3 relative;
  * the forward values of the differentiable route against the oracle (same 1e-4 as everywhere);
  * SVGP: d ELBO / d parameters against finite differences of the oracle's bound, and a short Adam run that must raise it.
"""
import numpy as np
import pytest
import torch

from oracle import gpsig_oracle as O
from util import assert_close, assert_levels_close, random_walks

pytestmark = pytest.mark.gpu


def _fd(f, x, eps=1e-6):
    """central finite differences of a scalar function of a numpy array"""
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


@pytest.mark.parametrize("shape,M", [((3, 20, 4, 17), 4), ((2, 33, 2, 40), 5), ((1, 5, 1, 3), 2), ((2, 130, 1, 129), 3)])
def test_sigkern_vjp_kernel_matches_the_oracle_vjp(shape, M):
    from gpsig_b200 import autodiff as AD
    rng = np.random.default_rng(0)
    Delta = rng.standard_normal(shape) / np.sqrt(shape[1])
    G = rng.standard_normal((M + 1, shape[0], shape[2]))
    D = torch.tensor(Delta, device="cuda", dtype=torch.float32, requires_grad=True)
    out = AD.SigKernFirstOrder.apply(D, M)
    assert_levels_close(out.detach().cpu().numpy(), O.signature_kern_first_order(Delta, M, difference=False), msg="forward")
    (out * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()
    ref = O.signature_kern_first_order_vjp(Delta, M, G)
    assert_close(D.grad.cpu().numpy(), ref, tol=1e-4, msg="dL/dDelta")


def test_sigkern_vjp_through_the_differencing_and_3d_input():
    from gpsig_b200 import autodiff as AD
    rng = np.random.default_rng(1)
    Mx = rng.standard_normal((5, 12, 12))
    G = rng.standard_normal((4, 5))
    T = torch.tensor(Mx, device="cuda", dtype=torch.float32, requires_grad=True)
    out = AD.sigkern_first_order(T, 3, difference=True)
    assert_levels_close(out.detach().cpu().numpy(), O.signature_kern_first_order(Mx, 3, difference=True), msg="forward 3-D")
    (out * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()
    ref = _fd(lambda A: float(np.sum(G * O.signature_kern_first_order(A, 3, difference=True))), Mx)
    assert_close(T.grad.cpu().numpy(), ref, tol=1e-3, msg="dL/dM")


@pytest.mark.parametrize("nz,n,Lh,M", [(3, 4, 15, 3), (2, 3, 40, 5), (1, 1, 2, 1)])
def test_tens_vs_seq_vjp_kernel_matches_finite_differences(nz, n, Lh, M):
    from gpsig_b200 import autodiff as AD
    rng = np.random.default_rng(2)
    T = M * (M + 1) // 2
    H = rng.standard_normal((T, nz, n, Lh)) / np.sqrt(Lh)
    G = rng.standard_normal((M + 1, nz, n))
    Ht = torch.tensor(H, device="cuda", dtype=torch.float32, requires_grad=True)
    out = AD.TensVsSeqFirstOrder.apply(Ht, M)
    fwd = lambda A: O.signature_kern_tens_vs_seq_first_order(A, M, difference=False)  # noqa: E731
    assert_levels_close(out.detach().cpu().numpy(), fwd(H), msg="forward")
    (out * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()
    ref = _fd(lambda A: float(np.sum(G * fwd(A))), H)
    assert_close(Ht.grad.cpu().numpy(), ref, tol=1e-3, msg="dL/dH")


def _setup(kind, L=20, d=3, M=3, n=6, nz=4, increments=True, seed=3, **kw):
    from gpsig_b200 import kernels
    rng = np.random.default_rng(seed)
    X = random_walks(n, L, d, seed).reshape(n, -1)
    T = M * (M + 1) // 2
    Z = 0.5 * rng.standard_normal((T, nz, 2, d) if increments else (T, nz, d))
    ls = np.array([0.9, 1.1, 1.4])[:d]
    var = 0.5 + rng.random(M + 1)
    cls = dict(linear=kernels.SignatureLinear, rbf=kernels.SignatureRBF, matern32=kernels.SignatureMatern32)[kind]
    k = cls(L * d, d, M, lengthscales=ls, variances=var, **kw)
    return k, X, Z, ls, var, rng


def _oracle(kind, L, d, M, ls, var, **kw):
    return O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=ls, variances=var, **kw)


@pytest.mark.parametrize("kind", ["rbf", "linear", "matern32"])
@pytest.mark.parametrize("norm", [True, False])
def test_K_tens_n_seq_covs_gradients(kind, norm):
    """configs[0]-sized problem: every output of the one call SVGP makes, every kind of parameter"""
    L, d, M = 20, 3, 3
    k, X, Z, ls, var, rng = _setup(kind, L, d, M, normalization=norm)
    wzz, wzx, wxx = rng.standard_normal((4, 4)), rng.standard_normal((4, 6)), rng.standard_normal(6)

    def loss_np(ls_, var_, Z_, X_):
        ko = _oracle(kind, L, d, M, ls_, var_, normalization=norm)
        Kzz, Kzx, Kxx = ko.K_tens_n_seq_covs(Z_, X_, increments=True)
        return float(np.sum(wzz * Kzz) + np.sum(wzx * Kzx) + np.sum(wxx * Kxx))

    k.set_trainable(("variances", "lengthscales"))
    Zt = torch.tensor(Z, device="cuda", dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(X, device="cuda", dtype=torch.float64, requires_grad=True)
    Kzz, Kzx, Kxx = k.K_tens_n_seq_covs(Zt, Xt, increments=True)
    ko = _oracle(kind, L, d, M, ls, var, normalization=norm)
    rzz, rzx, rxx = ko.K_tens_n_seq_covs(Z, X, increments=True)
    assert_close(Kzz.detach().cpu().numpy(), rzz, msg="Kzz")
    assert_close(Kzx.detach().cpu().numpy(), rzx, msg="Kzx")
    assert_close(Kxx.detach().cpu().numpy(), rxx, msg="Kxx")
    tt = lambda a: torch.tensor(a, device="cuda", dtype=Kzz.dtype)  # noqa: E731
    loss = (tt(wzz) * Kzz).sum() + (tt(wzx) * Kzx).sum() + (tt(wxx) * Kxx).sum()
    loss.backward()
    # parameters live in unconstrained space: chain rule through softplus for the comparison
    raw_ls, raw_var = k._raw["lengthscales"], k._raw["variances"]
    dls = raw_ls.grad.cpu().numpy() / torch.sigmoid(raw_ls.detach()).cpu().numpy()
    dvar = raw_var.grad.cpu().numpy() / torch.sigmoid(raw_var.detach()).cpu().numpy()
    assert_close(dls, _fd(lambda a: loss_np(a, var, Z, X), ls), tol=1e-3, msg="d/d lengthscales")
    assert_close(dvar, _fd(lambda a: loss_np(ls, a, Z, X), var), tol=1e-3, msg="d/d variances")
    assert_close(Zt.grad.cpu().numpy(), _fd(lambda a: loss_np(ls, var, a, X), Z), tol=1e-3, msg="d/dZ")
    assert_close(Xt.grad.cpu().numpy(), _fd(lambda a: loss_np(ls, var, Z, a), X), tol=1e-3, msg="d/dX")


@pytest.mark.parametrize("kind", ["rbf", "linear"])
def test_K_and_Kdiag_gradients_headline_tile(kind):
    """the headline tile shape L=128, d=8, M=5 on a few sequences: d/dX of sum(w * K(X, X2)) and of Kdiag"""
    from gpsig_b200 import kernels
    L, d, M = 128, 8, 5
    rng = np.random.default_rng(4)
    X, X2 = random_walks(3, L, d, 5).reshape(3, -1), random_walks(2, L, d, 6).reshape(2, -1)
    cls = dict(linear=kernels.SignatureLinear, rbf=kernels.SignatureRBF)[kind]
    ls = np.sqrt(8.0) if kind == "rbf" else 1.0
    k = cls(L * d, d, M, lengthscales=ls)
    ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=ls)
    w = rng.standard_normal((3, 2))
    Xt = torch.tensor(X, device="cuda", dtype=torch.float64, requires_grad=True)
    K = k.K(Xt, X2)
    assert_close(K.detach().cpu().numpy(), ko.K(X, X2), msg="K forward")
    (torch.tensor(w, device="cuda", dtype=K.dtype) * K).sum().backward()
    # directional finite differences (the full Jacobian has 3072 entries)
    for seed in range(3):
        v = np.random.default_rng(10 + seed).standard_normal(X.shape)
        eps = 1e-5
        fd = (np.sum(w * ko.K(X + eps * v, X2)) - np.sum(w * ko.K(X - eps * v, X2))) / (2 * eps)
        got = float(np.sum(Xt.grad.cpu().numpy() * v))
        assert abs(got - fd) <= 1e-3 * max(abs(fd), 1e-3), (kind, got, fd)


def test_svgp_elbo_gradient_and_training():
    from gpsig_b200 import models, inducing_variables as iv
    L, d, M, n, nz = 20, 3, 3, 12, 5
    k, X, Z, ls, var, rng = _setup("rbf", L, d, M, n=n, nz=nz)
    Y = (rng.standard_normal((n, 1)) > 0).astype(np.float64)
    q_mu = 0.3 * rng.standard_normal((nz, 1))
    q_sqrt = np.tril(0.2 * rng.standard_normal((1, nz, nz))) + np.eye(nz)[None]
    m = models.SVGP(X, Y, k, models.Bernoulli(), iv.InducingTensors(Z, M, increments=True), num_latent=1, q_mu=q_mu, q_sqrt=q_sqrt)
    ko = _oracle("rbf", L, d, M, ls, var)
    ref = O.svgp_elbo(ko, Z, X, Y, q_mu, q_sqrt, likelihood="bernoulli", increments=True)[0]
    assert abs(m.compute_log_likelihood() - ref) < 1e-4 * abs(ref)
    m.set_trainable()
    loss = m.training_loss()
    assert abs(-float(loss.item()) - ref) < 1e-4 * abs(ref)
    loss.backward()
    g_mu = -m.q_mu.grad.cpu().numpy()
    fd_mu = _fd(lambda a: O.svgp_elbo(ko, Z, X, Y, a, q_sqrt, likelihood="bernoulli", increments=True)[0], q_mu)
    assert_close(g_mu, fd_mu, tol=1e-3, msg="dELBO/dq_mu")
    g_Z = -m.feature.Z.grad.cpu().numpy()
    fd_Z = _fd(lambda a: O.svgp_elbo(ko, a, X, Y, q_mu, q_sqrt, likelihood="bernoulli", increments=True)[0], Z)
    assert_close(g_Z, fd_Z, tol=1e-3, msg="dELBO/dZ")
    raw = k._raw["lengthscales"]
    g_ls = -raw.grad.cpu().numpy() / torch.sigmoid(raw.detach()).cpu().numpy()
    fd_ls = _fd(lambda a: O.svgp_elbo(_oracle("rbf", L, d, M, a, var), Z, X, Y, q_mu, q_sqrt, likelihood="bernoulli",
                                      increments=True)[0], ls)
    assert_close(g_ls, fd_ls, tol=1e-3, msg="dELBO/dlengthscales")
    hist = m.optimize(iterations=30, lr=5e-2)
    assert hist[-1] > hist[0] + 1e-3, hist
    assert np.isfinite(hist).all()
    # the trained parameters are visible to the fast (non-differentiable) path
    with torch.no_grad():
        e_fast = m.compute_log_likelihood()
    e_slow = -float(m.training_loss().item())
    assert abs(e_fast - e_slow) < 1e-3 * abs(e_slow)


def test_multiclass_and_mean_function_train():
    from gpsig_b200 import models, inducing_variables as iv
    L, d, M, n, nz, C = 16, 2, 3, 15, 4, 3
    k, X, Z, ls, var, rng = _setup("rbf", L, d, M, n=n, nz=nz)
    Y = rng.integers(0, C, size=(n, 1)).astype(np.float64)
    m = models.SVGP(X, Y, k, models.MultiClass(C), iv.InducingTensors(Z, M, increments=True), num_latent=C,
                    mean_function=models.Constant(np.zeros(C)))
    e0 = m.compute_log_likelihood()
    assert np.isfinite(e0)
    # with q = prior and zero mean every class is equally likely: E log p = n [ p log(1-eps) + (1-p) log(eps/(C-1)) ], p = 1/C
    p, eps = 1.0 / C, 1e-3
    assert abs(e0 - n * (p * np.log(1 - eps) + (1 - p) * np.log(eps / (C - 1)))) < 0.05 * abs(e0)
    hist = m.optimize(iterations=25, lr=5e-2)
    assert hist[-1] > hist[0]
    probs = m.likelihood.predict_mean(*m.predict_f(X)).detach().cpu().numpy()
    assert probs.shape == (n, C) and np.allclose(probs.sum(1), 1.0, atol=2e-2)


def test_blockwise_checkpointed_gram_gives_the_same_values_and_gradients():
    """settings.autodiff_gram_budget_bytes = 0 forces the blockwise, activation-checkpointed evaluation of the Gram on both
    differentiable routes (K and K_tens_n_seq_covs): values and gradients must agree with the resident evaluation."""
    from gpsig_b200 import settings
    L, d, M = 20, 3, 3
    results = []
    old = settings.autodiff_gram_budget_bytes
    try:
        for budget in (old, 0):
            settings.autodiff_gram_budget_bytes = budget
            k, X, Z, ls, var, rng = _setup("rbf", L, d, M, n=9, nz=5)
            k.set_trainable(("variances", "lengthscales"))
            Zt = torch.tensor(Z, device="cuda", dtype=torch.float64, requires_grad=True)
            Xt = torch.tensor(X, device="cuda", dtype=torch.float64, requires_grad=True)
            Kzz, Kzx, Kxx = k.K_tens_n_seq_covs(Zt, Xt, increments=True)
            Kxx_full = k.K(Xt)
            w = torch.linspace(-1.0, 1.0, Kzx.numel(), device="cuda", dtype=Kzx.dtype).reshape(Kzx.shape)
            loss = (w * Kzx).sum() + Kzz.sum() + (Kxx * Kxx).sum() + (Kxx_full * Kxx_full).sum()
            loss.backward()
            results.append([loss.detach().cpu().numpy(), Zt.grad.cpu().numpy(), Xt.grad.cpu().numpy(),
                            k._raw["lengthscales"].grad.cpu().numpy(), k._raw["variances"].grad.cpu().numpy()])
    finally:
        settings.autodiff_gram_budget_bytes = old
    for a, b, name in zip(results[0], results[1], ("loss", "dZ", "dX", "dlengthscales", "dvariances")):
        assert_close(b, a, tol=1e-5, msg=name)
