"""
SVGP -- drop-in consumer of Kuu_Kuf_Kff, mirroring gpsig/models.py:13-73 (forward pass: ELBO and predictive moments).

The covariances come from the CUDA path; the downstream dense algebra (Cholesky of Kzz, triangular solves, q_sqrt
matmuls -- gpflow's base_conditional / gauss_kl, GPflow 1.5.1, not part of /root/reference) is cuSOLVER/cuBLAS through
torch.linalg, where tensor cores are the right tool.  Gradients / training are outside this round's scope
(SURVEY.md 8f rank 1).
"""
import math

import numpy as np
import torch

from . import settings
from .inducing_variables import InducingTensors, InducingSequences, Kuu_Kuf_Kff


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """gpflow/conditionals.py base_conditional (1.5.1), called at models.py:66.  f (Z, R); q_sqrt (R, Z, Z) or (Z, R)."""
    R = f.shape[1]
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    if full_cov:
        fvar = (Knn - A.transpose(0, 1) @ A)[None].expand(R, -1, -1)
    else:
        fvar = (Knn - torch.sum(A * A, 0))[None].expand(R, -1)
    if not white:
        A = torch.linalg.solve_triangular(Lm.transpose(0, 1), A, upper=True)
    fmean = A.transpose(0, 1) @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt.transpose(0, 1)[:, :, None]
        else:
            LTA = torch.matmul(torch.tril(q_sqrt).transpose(-1, -2), A[None])
        if full_cov:
            fvar = fvar + torch.matmul(LTA.transpose(-1, -2), LTA)
        else:
            fvar = fvar + torch.sum(LTA * LTA, 1)
    if not full_cov:
        fvar = fvar.transpose(0, 1)
    return fmean, fvar


def gauss_kl(q_mu, q_sqrt, K=None):
    """gpflow/kullback_leiblers.py gauss_kl (1.5.1), called at models.py:49,52."""
    Zn, R = q_mu.shape
    white = K is None
    if white:
        alpha = q_mu
    else:
        Lp = torch.linalg.cholesky(K)
        alpha = torch.linalg.solve_triangular(Lp, q_mu, upper=False)
    if q_sqrt.dim() == 2:
        Lq_diag = q_sqrt
        Lq_full = torch.diag_embed(q_sqrt.transpose(0, 1))
    else:
        Lq_full = torch.tril(q_sqrt)
        Lq_diag = torch.diagonal(Lq_full, dim1=-2, dim2=-1)
    twoKL = torch.sum(alpha * alpha) - R * Zn - torch.sum(torch.log(Lq_diag * Lq_diag))
    if white:
        twoKL = twoKL + torch.sum(q_sqrt * q_sqrt if q_sqrt.dim() == 2 else Lq_full * Lq_full)
    else:
        LpiLq = torch.linalg.solve_triangular(Lp[None].expand(R, -1, -1), Lq_full, upper=False)
        twoKL = twoKL + torch.sum(LpiLq * LpiLq) + R * torch.sum(torch.log(torch.diagonal(Lp) ** 2))
    return 0.5 * twoKL


class Gaussian:
    """gpflow.likelihoods.Gaussian (variational_expectations only)."""

    def __init__(self, variance=1.0):
        self.variance = float(variance)

    def variational_expectations(self, Fmu, Fvar, Y):
        return -0.5 * math.log(2 * math.pi) - 0.5 * math.log(self.variance) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / self.variance


class Bernoulli:
    """gpflow.likelihoods.Bernoulli with the probit link (inv_probit jitter 1e-3), 20-point Gauss-Hermite quadrature."""

    def __init__(self, num_gauss_hermite_points=20):
        x, w = np.polynomial.hermite.hermgauss(num_gauss_hermite_points)
        self._x, self._w = x, w / np.sqrt(np.pi)

    def variational_expectations(self, Fmu, Fvar, Y):
        x = torch.as_tensor(self._x, device=Fmu.device, dtype=Fmu.dtype)
        w = torch.as_tensor(self._w, device=Fmu.device, dtype=Fmu.dtype)
        X = Fmu[..., None] + torch.sqrt(2.0 * Fvar[..., None]) * x
        p = 0.5 * (1.0 + torch.erf(X / math.sqrt(2.0))) * (1 - 2e-3) + 1e-3
        logp = torch.where(Y[..., None] == 1, torch.log(p), torch.log(1 - p))
        return torch.sum(logp * w, dim=-1)


class SVGP:
    """models.py:13-73.  Holds X, Y, kern, likelihood, feat, q_mu (Z, R), q_sqrt (R, Z, Z) / (Z, R)."""

    def __init__(self, X, Y, kern, likelihood, feat, mean_function=None, num_latent=None, q_diag=False, whiten=True,
                 minibatch_size=None, num_data=None, q_mu=None, q_sqrt=None, shuffle=True, **kwargs):
        if not isinstance(feat, InducingTensors) and not isinstance(feat, InducingSequences):
            raise ValueError('feat must be of type either InducingTensors or InducingSequences')
        if mean_function is not None:
            raise NotImplementedError("only the zero mean function is supported")
        num_inducing = len(feat)
        self.X, self.Y = X, Y
        self.kern, self.likelihood, self.feature = kern, likelihood, feat
        self.num_latent = num_latent or Y.shape[1]
        self.num_data = num_data or X.shape[0]
        self.q_diag, self.whiten = q_diag, whiten
        self.minibatch_size = minibatch_size
        self._rng = np.random.RandomState(0)                                                      # Minibatch(seed=0)
        self.shuffle = shuffle
        # gpflow SVGP._init_variational_parameters
        self.q_mu = np.zeros((num_inducing, self.num_latent)) if q_mu is None else np.asarray(q_mu, dtype=np.float64)
        if q_sqrt is None:
            self.q_sqrt = (np.ones((num_inducing, self.num_latent)) if q_diag
                           else np.tile(np.eye(num_inducing)[None], [self.num_latent, 1, 1]))
        else:
            self.q_sqrt = np.asarray(q_sqrt, dtype=np.float64)

    def _dev(self, a, dev, dtype=torch.float32):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=dtype)
        return torch.as_tensor(np.asarray(a)).to(device=dev, dtype=dtype)

    def _build_predict(self, X_new, full_cov=False, full_output_cov=False, return_Kzz=False):
        """models.py:61-73."""
        Kzz, Kzx, Kxx = Kuu_Kuf_Kff(self.feature, self.kern, X_new, jitter=settings.jitter, full_f_cov=full_cov)
        dev = Kzz.device
        q_mu = self._dev(self.q_mu, dev)
        q_sqrt = self._dev(self.q_sqrt, dev)
        if q_sqrt.dim() == 3:
            q_sqrt = torch.tril(q_sqrt)                                                            # matrix_band_part(-1, 0)
        f_mean, f_var = base_conditional(Kzx, Kzz, Kxx, q_mu, full_cov=full_cov, q_sqrt=q_sqrt, white=self.whiten)
        if return_Kzz:
            return f_mean, f_var, Kzz
        return f_mean, f_var

    def _batch(self):
        if self.minibatch_size is None or self.minibatch_size >= self.X.shape[0]:
            return self.X, self.Y
        idx = self._rng.permutation(self.X.shape[0])[:self.minibatch_size] if self.shuffle else np.arange(self.minibatch_size)
        return self.X[idx], self.Y[idx]

    def _build_likelihood(self, X=None, Y=None):
        """models.py:39-59: ELBO = sum(var_exp) * num_data / batch - KL."""
        if X is None:
            X, Y = self._batch()
        num_samples = X.shape[0]
        if self.whiten:
            f_mean, f_var = self._build_predict(X)
            Kzz = None
        else:
            f_mean, f_var, Kzz = self._build_predict(X, return_Kzz=True)
        dev = f_mean.device
        q_sqrt = self._dev(self.q_sqrt, dev)
        KL = gauss_kl(self._dev(self.q_mu, dev), torch.tril(q_sqrt) if q_sqrt.dim() == 3 else q_sqrt, K=Kzz)
        var_exp = self.likelihood.variational_expectations(f_mean, f_var, self._dev(Y, dev))
        scale = float(self.num_data) / float(num_samples)
        return torch.sum(var_exp) * scale - KL

    def compute_log_likelihood(self, X=None, Y=None):
        return float(self._build_likelihood(X, Y).item())

    def predict_f(self, X_new, full_cov=False):
        return self._build_predict(X_new, full_cov=full_cov)
