"""
Independent pin of the GPflow-1.5.1 algebra restated in gpsig_b200/models.py (base_conditional, gauss_kl, the likelihoods):
dense joint-Gaussian conditioning and textbook KL in float64 NumPy, written here from the definitions -- no code shared with
oracle/ or with the package.  (GPflow is not part of /root/reference: VERDICT r1 asked for a pin beyond self-consistency.)

  u ~ N(0, Kmm),  f | u ~ N(Knm Kmm^-1 u, Knn - Knm Kmm^-1 Kmn),  q(u) = N(m, S):
      q(f) = N(Knm Kmm^-1 m,  Knn - Knm Kmm^-1 (Kmm - S) Kmm^-1 Kmn)
  whitened: u = L v, q(v) = N(m_v, S_v)  =>  m = L m_v, S = L S_v L^T.
  KL[N(m, S) || N(0, K)] = 1/2 (tr(K^-1 S) + m^T K^-1 m - M + log det K - log det S).
"""
import numpy as np
import pytest
import torch

from gpsig_b200 import models


def _spd(rng, n, scale=1.0):
    A = rng.standard_normal((n, n))
    return scale * (A @ A.T / n + 0.3 * np.eye(n))


def _setup(seed, M=7, N=5, R=2):
    rng = np.random.default_rng(seed)
    J = _spd(rng, M + N)                       # joint prior covariance of (u, f)
    Kmm, Kmn, Knn = J[:M, :M], J[:M, M:], J[M:, M:]
    q_mu = rng.standard_normal((M, R))
    q_sqrt = np.tril(rng.standard_normal((R, M, M))) + 2.0 * np.eye(M)[None]
    return rng, Kmm, Kmn, Knn, q_mu, q_sqrt


def _dense_posterior(Kmm, Kmn, Knn, m, S):
    A = np.linalg.solve(Kmm, Kmn)              # Kmm^-1 Kmn
    mean = A.T @ m
    cov = Knn - A.T @ (Kmm - S) @ A
    return mean, cov


@pytest.mark.parametrize("white", [True, False])
@pytest.mark.parametrize("full_cov", [True, False])
def test_base_conditional_against_dense_joint_gaussian(white, full_cov):
    rng, Kmm, Kmn, Knn, q_mu, q_sqrt = _setup(0)
    L = np.linalg.cholesky(Kmm)
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    fm, fv = models.base_conditional(t(Kmn), t(Kmm), t(Knn if full_cov else np.diag(Knn)), t(q_mu), full_cov=full_cov,
                                     q_sqrt=t(q_sqrt), white=white)
    for r in range(q_mu.shape[1]):
        S_r = q_sqrt[r] @ q_sqrt[r].T
        m_r = q_mu[:, r]
        if white:
            m_r, S_r = L @ m_r, L @ S_r @ L.T
        mean, cov = _dense_posterior(Kmm, Kmn, Knn, m_r, S_r)
        np.testing.assert_allclose(fm[:, r].numpy(), mean, rtol=1e-9, atol=1e-10)
        if full_cov:
            np.testing.assert_allclose(fv[r].numpy(), cov, rtol=1e-9, atol=1e-10)
        else:
            np.testing.assert_allclose(fv[:, r].numpy(), np.diag(cov), rtol=1e-9, atol=1e-10)


def test_base_conditional_diagonal_q_sqrt():
    rng, Kmm, Kmn, Knn, q_mu, _ = _setup(1)
    q_diag = 0.5 + rng.random(q_mu.shape)      # (M, R): standard deviations
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    fm, fv = models.base_conditional(t(Kmn), t(Kmm), t(np.diag(Knn)), t(q_mu), full_cov=False, q_sqrt=t(q_diag), white=False)
    for r in range(q_mu.shape[1]):
        mean, cov = _dense_posterior(Kmm, Kmn, Knn, q_mu[:, r], np.diag(q_diag[:, r] ** 2))
        np.testing.assert_allclose(fm[:, r].numpy(), mean, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(fv[:, r].numpy(), np.diag(cov), rtol=1e-9, atol=1e-10)


def _kl_dense(m, S, K):
    Mn = len(m)
    Ki = np.linalg.inv(K)
    return 0.5 * (np.trace(Ki @ S) + m @ Ki @ m - Mn + np.linalg.slogdet(K)[1] - np.linalg.slogdet(S)[1])


@pytest.mark.parametrize("white", [True, False])
def test_gauss_kl_against_the_textbook_formula(white):
    rng, Kmm, _, _, q_mu, q_sqrt = _setup(2)
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    got = float(models.gauss_kl(t(q_mu), t(q_sqrt), K=None if white else t(Kmm)))
    K = np.eye(len(Kmm)) if white else Kmm
    ref = sum(_kl_dense(q_mu[:, r], q_sqrt[r] @ q_sqrt[r].T, K) for r in range(q_mu.shape[1]))
    assert abs(got - ref) < 1e-9 * abs(ref)
    # diagonal q_sqrt
    q_diag = 0.5 + rng.random(q_mu.shape)
    got = float(models.gauss_kl(t(q_mu), t(q_diag), K=None if white else t(Kmm)))
    ref = sum(_kl_dense(q_mu[:, r], np.diag(q_diag[:, r] ** 2), K) for r in range(q_mu.shape[1]))
    assert abs(got - ref) < 1e-9 * abs(ref)


def test_likelihood_expectations_against_brute_force_quadrature():
    """E_{f ~ N(mu, var)} log p(y | f) by dense trapezoidal integration"""
    rng = np.random.default_rng(3)
    mu, var = rng.standard_normal((6, 1)), 0.2 + rng.random((6, 1))
    y = (rng.random((6, 1)) > 0.5).astype(np.float64)
    f = np.linspace(-12, 12, 200001)
    from math import erf, sqrt, pi, log
    Phi = 0.5 * (1 + np.vectorize(erf)(f / sqrt(2.0)))
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    got_b = models.Bernoulli().variational_expectations(t(mu), t(var), t(y)).numpy()
    got_g = models.Gaussian(0.7).variational_expectations(t(mu), t(var), t(y)).numpy()
    for i in range(6):
        dens = np.exp(-0.5 * (f - mu[i, 0]) ** 2 / var[i, 0]) / np.sqrt(2 * pi * var[i, 0])
        p = Phi * (1 - 2e-3) + 1e-3                                   # gpflow inv_probit jitter
        logp = np.log(p) if y[i, 0] == 1 else np.log(1 - p)
        assert abs(np.trapezoid(dens * logp, f) - got_b[i, 0]) < 2e-5      # 20-point Gauss-Hermite (gpflow default) vs dense integration
        lg = -0.5 * log(2 * pi * 0.7) - 0.5 * (y[i, 0] - f) ** 2 / 0.7
        assert abs(np.trapezoid(dens * lg, f) - got_g[i, 0]) < 1e-8


def test_multiclass_robustmax_against_monte_carlo():
    rng = np.random.default_rng(4)
    N, K = 4, 3
    mu, var = rng.standard_normal((N, K)), 0.3 + rng.random((N, K))
    y = rng.integers(0, K, size=(N, 1))
    lik = models.MultiClass(K)
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    p = lik.prob_is_largest(t(y.astype(np.float64)), t(mu), t(var)).numpy()
    S = 400000
    f = mu[None] + np.sqrt(var)[None] * rng.standard_normal((S, N, K))
    mc = (np.argmax(f, axis=2) == y[:, 0][None]).mean(0)
    np.testing.assert_allclose(p, mc, atol=5e-3)
    ve = lik.variational_expectations(t(mu), t(var), t(y.astype(np.float64))).numpy()[:, 0]
    np.testing.assert_allclose(ve, p * np.log(1 - 1e-3) + (1 - p) * np.log(1e-3 / (K - 1)), rtol=1e-12)
