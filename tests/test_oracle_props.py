"""
Property checks of the oracle that replace the reference's esig notebook (notebooks/signature_kernel.ipynb:52-310)
and cover what the reference never pins (order=1, diag path, gpflow formulas).  CPU only.
"""
import numpy as np
import pytest

from oracle import gpsig_oracle as O


def _walk(rng, n, L, d):
    return np.cumsum(rng.standard_normal((n, L, d)), axis=1) / np.sqrt(L)


def test_order_M_linear_equals_signature_inner_products():
    """notebook :52-140 (esig -> Chen identity): SignatureLinear(order=M, normalization=False) == <S(x), S(y)>."""
    rng = np.random.default_rng(0)
    n, L, d, M = 6, 9, 3, 4
    X = rng.standard_normal((n, L, d))
    sigs = np.stack([O.chen_signature(x, M) for x in X])
    k = O.SignatureKernelOracle("linear", L * d, d, M, order=M, normalization=False, lengthscales=None)
    K = k.K(X.reshape(n, -1))
    np.testing.assert_allclose(K, sigs @ sigs.T, rtol=1e-10, atol=1e-10)


def test_tensor_vs_seq_and_tensor_vs_tensor_equal_explicit_tensors():
    """notebook :167-310: K_tens_vs_seq == tens @ sigs.T, K_tens == tens @ tens.T for order=M linear."""
    rng = np.random.default_rng(1)
    n, L, d, M, nz = 5, 8, 3, 4, 7
    X = rng.standard_normal((n, L, d))
    Z = rng.standard_normal((M * (M + 1) // 2, nz, d))
    sigs = np.stack([O.chen_signature(x, M) for x in X])
    tens = O.rank1_tensors(Z, M)
    k = O.SignatureKernelOracle("linear", L * d, d, M, order=M, normalization=False, lengthscales=None)
    np.testing.assert_allclose(k.K_tens_vs_seq(Z, X.reshape(n, -1)), tens @ sigs.T, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(k.K_tens(Z), tens @ tens.T, rtol=1e-10, atol=1e-10)


def test_first_order_equals_bruteforce_enumeration():
    rng = np.random.default_rng(2)
    D = rng.standard_normal((5, 6))
    got = O.signature_kern_first_order(D[None, :, None, :], 3, difference=False)[:, 0, 0]
    np.testing.assert_allclose(got, O.brute_force_first_order(D, 3), rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_diag_path_equals_diagonal_of_full_path(order):
    rng = np.random.default_rng(3)
    X = _walk(rng, 4, 7, 2)
    k = O.SignatureKernelOracle("rbf", 14, 2, 4, order=order, normalization=False)
    full = k._K_seq(X)
    diag = k._K_seq_diag(X)
    np.testing.assert_allclose(diag, np.diagonal(full, axis1=1, axis2=2), rtol=1e-10, atol=1e-12)


def test_symmetric_vs_rectangular_and_normalised_diag():
    rng = np.random.default_rng(4)
    X = _walk(rng, 5, 8, 3).reshape(5, -1)
    k = O.SignatureKernelOracle("rbf", 24, 3, 3, variances=[1.0, 0.5, 2.0, 0.25], sigma=1.3)
    Ks, Kr = k.K(X), k.K(X, X)
    off = ~np.eye(5, dtype=bool)
    # quirk Q7: symmetric adds jitter on the diagonal only, rectangular on both diag vectors -> agree to ~jitter
    np.testing.assert_allclose(Ks[off], Kr[off], rtol=1e-4)
    np.testing.assert_allclose(np.diag(Ks), k.Kdiag(X), rtol=1e-12)


def test_higher_order_with_order_1_is_first_order():
    rng = np.random.default_rng(5)
    M = rng.standard_normal((3, 6, 2, 5))
    np.testing.assert_allclose(O.signature_kern_higher_order(M, 4, order=1), O.signature_kern_first_order(M, 4),
                               rtol=1e-12, atol=1e-13)
    Mt = rng.standard_normal((10, 3, 2, 6))
    np.testing.assert_allclose(O.signature_kern_tens_vs_seq_higher_order(Mt, 4, order=1),
                               O.signature_kern_tens_vs_seq_first_order(Mt, 4), rtol=1e-12, atol=1e-13)


def test_base_conditional_and_gauss_kl_against_dense_gaussian_algebra():
    """gpflow formulas (parity unpinned): compare with the textbook dense expressions."""
    rng = np.random.default_rng(6)
    Zn, N, R = 5, 7, 2
    A = rng.standard_normal((Zn + N, Zn + N))
    Kfull = A @ A.T + 0.5 * np.eye(Zn + N)
    Kmm, Kmn, Knn = Kfull[:Zn, :Zn], Kfull[:Zn, Zn:], np.diag(Kfull[Zn:, Zn:])
    q_mu = rng.standard_normal((Zn, R))
    q_sqrt = np.tril(rng.standard_normal((R, Zn, Zn))) + 2 * np.eye(Zn)[None]
    fm, fv = O.base_conditional(Kmn, Kmm, Knn, q_mu, q_sqrt=q_sqrt, white=False)
    Ki = np.linalg.inv(Kmm)
    np.testing.assert_allclose(fm, Kmn.T @ Ki @ q_mu, rtol=1e-9)
    for r in range(R):
        S = q_sqrt[r] @ q_sqrt[r].T
        want = Knn - np.sum(Kmn * (Ki @ Kmn), 0) + np.sum(Kmn * (Ki @ S @ Ki @ Kmn), 0)
        np.testing.assert_allclose(fv[:, r], want, rtol=1e-9)
    # KL[N(q_mu, S) || N(0, Kmm)] summed over outputs
    want = 0.0
    for r in range(R):
        S = q_sqrt[r] @ q_sqrt[r].T
        want += 0.5 * (np.trace(Ki @ S) + q_mu[:, r] @ Ki @ q_mu[:, r] - Zn + np.linalg.slogdet(Kmm)[1]
                       - np.linalg.slogdet(S)[1])
    np.testing.assert_allclose(O.gauss_kl(q_mu, q_sqrt, Kmm), want, rtol=1e-9)
    # whitened: p = N(0, I)
    want = sum(0.5 * (np.trace(q_sqrt[r] @ q_sqrt[r].T) + q_mu[:, r] @ q_mu[:, r] - Zn
                      - np.linalg.slogdet(q_sqrt[r] @ q_sqrt[r].T)[1]) for r in range(R))
    np.testing.assert_allclose(O.gauss_kl(q_mu, q_sqrt), want, rtol=1e-9)


def test_first_order_vjp_matches_finite_differences():
    """groundwork for the backward pass: reverse-mode recursion (suffix sums) against central differences."""
    rng = np.random.default_rng(17)
    n1, r, n2, c, M = 2, 5, 3, 4, 4
    Delta = 0.3 * rng.standard_normal((n1, r, n2, c))
    G = rng.standard_normal((M + 1, n1, n2))
    loss = lambda D: float(np.sum(G * O.signature_kern_first_order(D, M, difference=False)))  # noqa: E731
    grad = O.signature_kern_first_order_vjp(Delta, M, G)
    num = np.zeros_like(Delta)
    h = 1e-6
    it = np.nditer(Delta, flags=["multi_index"])
    for _ in it:
        idx = it.multi_index
        Dp, Dm = Delta.copy(), Delta.copy()
        Dp[idx] += h
        Dm[idx] -= h
        num[idx] = (loss(Dp) - loss(Dm)) / (2 * h)
    np.testing.assert_allclose(grad, num, rtol=1e-6, atol=1e-8)
