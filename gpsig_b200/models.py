"""
SVGP -- drop-in consumer of Kuu_Kuf_Kff, mirroring gpsig/models.py:13-73: ELBO, predictive moments, and training.

The covariances come from the CUDA path; the downstream dense algebra (Cholesky of Kzz, triangular solves, q_sqrt
matmuls -- gpflow's base_conditional / gauss_kl, GPflow 1.5.1, not part of /root/reference) is cuSOLVER/cuBLAS through
torch.linalg in float64 like the reference (an M x M Cholesky with jitter 1e-6 has no business in fp32).

Training (the reference: gpsig/training.py:140-203 driving TF optimisers over the graph of models.py:39-59): call
`set_trainable()`; every parameter becomes a torch leaf in unconstrained space and `training_loss()` is differentiable --
the covariances then take the route of autodiff.py, whose two recursions have hand-written CUDA backward kernels.
"""
import math

import numpy as np
import torch

from . import settings
from .inducing_variables import InducingTensors, InducingSequences, Kuu_Kuf_Kff


_pending_info = []


def _cholesky(K):
    """torch.linalg.cholesky raises on failure, which costs a device synchronisation in the middle of a step (the host
    cannot queue the rest of the step behind the Kuf kernel).  On the device the status word is kept instead and checked
    where the step synchronises anyway (check_cholesky: compute_log_likelihood, optimize, predict_f)."""
    if not K.is_cuda:
        return torch.linalg.cholesky(K)
    L, info = torch.linalg.cholesky_ex(K)
    _pending_info.append(info)
    del _pending_info[:-32]
    return L


def check_cholesky():
    """Raise if a Cholesky factorisation queued since the last check failed (synchronises)."""
    bad = any(int(i.max()) != 0 for i in _pending_info)
    _pending_info.clear()
    if bad:
        raise torch.linalg.LinAlgError("Cholesky factorisation failed: the covariance matrix is not positive definite")


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """gpflow/conditionals.py base_conditional (1.5.1), called at models.py:66.  f (Z, R); q_sqrt (R, Z, Z) or (Z, R)."""
    R = f.shape[1]
    Lm = _cholesky(Kmm)
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
        Lp = _cholesky(K)
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
        self._raw = None

    def set_trainable(self, dev):
        from .autodiff import inv_softplus
        self._raw = torch.tensor(float(inv_softplus(self.variance)), dtype=torch.float64, device=dev, requires_grad=True)

    def parameters(self):
        return [] if self._raw is None else [self._raw]

    def _var(self, like):
        if self._raw is None:
            return torch.as_tensor(self.variance, device=like.device, dtype=like.dtype)
        from .autodiff import softplus
        return softplus(self._raw).to(like.device, like.dtype)

    def variational_expectations(self, Fmu, Fvar, Y):
        v = self._var(Fmu)
        return -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(v) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / v


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


class MultiClass:
    """gpflow.likelihoods.MultiClass with the RobustMax inverse link (1.5.1; benchmarks/models/train_gpsig.py:64 uses it):
    p = P(f_y is the largest latent) by Gauss-Hermite quadrature over f_y, the other latents integrated in closed form;
    E_q log p(y | f) = p log(1 - eps) + (1 - p) log(eps / (K - 1)).  Y holds class indices (N, 1)."""

    def __init__(self, num_classes, epsilon=1e-3, num_gauss_hermite_points=20):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)
        x, w = np.polynomial.hermite.hermgauss(num_gauss_hermite_points)
        self._x, self._w = x, w / np.sqrt(np.pi)

    def prob_is_largest(self, Y, mu, var):
        K = self.num_classes
        x = torch.as_tensor(self._x, device=mu.device, dtype=mu.dtype)
        w = torch.as_tensor(self._w, device=mu.device, dtype=mu.dtype)
        oh = torch.nn.functional.one_hot(Y.reshape(-1).to(torch.int64), K).to(mu.dtype)          # (N, K)
        mu_sel = torch.sum(oh * mu, dim=1, keepdim=True)
        var_sel = torch.sum(oh * var, dim=1, keepdim=True)
        X = mu_sel + torch.sqrt(torch.clamp(2.0 * var_sel, 1e-10, float("inf"))) * x[None, :]      # (N, H)
        dist = (X[:, None, :] - mu[:, :, None]) / torch.sqrt(torch.clamp(var, 1e-10, float("inf")))[:, :, None]
        cdfs = 0.5 * (1.0 + torch.erf(dist / math.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        cdfs = cdfs * (1.0 - oh[:, :, None]) + oh[:, :, None]                                       # the selected latent: factor 1
        return torch.prod(cdfs, dim=1) @ w                                                          # (N,)

    def variational_expectations(self, Fmu, Fvar, Y):
        p = self.prob_is_largest(Y, Fmu, Fvar)
        ve = p * math.log(1.0 - self.epsilon) + (1.0 - p) * math.log(self.epsilon / (self.num_classes - 1.0))
        return ve[:, None]

    def predict_mean(self, Fmu, Fvar):
        """class probabilities (N, K) (gpflow MultiClass.predict_mean_and_var's mean)"""
        N, K = Fmu.shape
        cols = []
        for k in range(K):
            yk = torch.full((N, 1), k, device=Fmu.device, dtype=torch.int64)
            p = self.prob_is_largest(yk, Fmu, Fvar)
            cols.append(p * (1.0 - self.epsilon) + (1.0 - p) * self.epsilon / (K - 1.0))
        return torch.stack(cols, dim=1)


class Zero:
    """gpflow.mean_functions.Zero"""

    def __call__(self, X):
        return 0.0

    def parameters(self):
        return []


class Constant:
    """gpflow.mean_functions.Constant: c broadcast over the rows of X"""

    def __init__(self, c):
        self.c = torch.as_tensor(np.atleast_1d(np.asarray(c, dtype=np.float64)))

    def __call__(self, X):
        return self.c.to(X.device)[None, :].expand(X.shape[0], -1)

    def parameters(self):
        return [self.c] if self.c.requires_grad else []


class Linear:
    """gpflow.mean_functions.Linear: X A + b on the flattened (N, L d) inputs"""

    def __init__(self, A, b):
        self.A = torch.as_tensor(np.asarray(A, dtype=np.float64))
        self.b = torch.as_tensor(np.atleast_1d(np.asarray(b, dtype=np.float64)))

    def __call__(self, X):
        return X.to(torch.float64) @ self.A.to(X.device) + self.b.to(X.device)

    def parameters(self):
        return [t for t in (self.A, self.b) if t.requires_grad]


class SVGP:
    """models.py:13-73.  Holds X, Y, kern, likelihood, feat, q_mu (Z, R), q_sqrt (R, Z, Z) / (Z, R)."""

    def __init__(self, X, Y, kern, likelihood, feat, mean_function=None, num_latent=None, q_diag=False, whiten=True,
                 minibatch_size=None, num_data=None, q_mu=None, q_sqrt=None, shuffle=True, **kwargs):
        if not isinstance(feat, InducingTensors) and not isinstance(feat, InducingSequences):
            raise ValueError('feat must be of type either InducingTensors or InducingSequences')
        self.mean_function = mean_function if mean_function is not None else Zero()               # gpflow GPModel: Zero()
        num_inducing = len(feat)
        self._trainable = False
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

    def _dev(self, a, dev, dtype=torch.float64):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=dtype)
        return torch.as_tensor(np.asarray(a)).to(device=dev, dtype=dtype)

    def _build_predict(self, X_new, full_cov=False, full_output_cov=False, return_Kzz=False, return_q=False):
        """models.py:61-73.  The covariances arrive in fp32 from the device path; the conditional runs in float64."""
        Kzz, Kzx, Kxx = Kuu_Kuf_Kff(self.feature, self.kern, X_new, jitter=settings.jitter, full_f_cov=full_cov)
        dev = Kzz.device
        Kzz, Kzx, Kxx = Kzz.to(torch.float64), Kzx.to(torch.float64), Kxx.to(torch.float64)
        q_mu = self._dev(self.q_mu, dev)
        q_sqrt = self._dev(self.q_sqrt, dev)
        if q_sqrt.dim() == 3:
            q_sqrt = torch.tril(q_sqrt)                                                            # matrix_band_part(-1, 0)
        f_mean, f_var = base_conditional(Kzx, Kzz, Kxx, q_mu, full_cov=full_cov, q_sqrt=q_sqrt, white=self.whiten)
        if not isinstance(self.mean_function, Zero):                                               # models.py:67
            f_mean = f_mean + self.mean_function(self._dev(X_new, dev).reshape(X_new.shape[0], -1))
        if return_q:   # the bound needs q_mu / q_sqrt on the device again (KL): uploaded once per step, not twice
            return f_mean, f_var, (Kzz if return_Kzz else None), q_mu, q_sqrt
        if return_Kzz:
            return f_mean, f_var, Kzz
        return f_mean, f_var

    # ---- training (gpsig/training.py:140-203: the reference optimises the same bound with TF optimisers) ----
    def set_trainable(self, kernel=("variances", "sigma", "lengthscales"), inducing=True, variational=True, likelihood=True,
                      device=None):
        """Make the model's parameters torch leaves: kernel parameters (unconstrained, kern.set_trainable), the inducing
        tensors / sequences Z, q_mu and q_sqrt, and the Gaussian likelihood's variance."""
        dev = torch.device(device) if device is not None else self.kern._dev()
        if kernel:
            self.kern.set_trainable(kernel, device=dev)
        if inducing:
            self.feature.Z = self._dev(self.feature.Z, dev).clone().requires_grad_(True)
        if variational:
            self.q_mu = self._dev(self.q_mu, dev).clone().requires_grad_(True)
            self.q_sqrt = self._dev(self.q_sqrt, dev).clone().requires_grad_(True)
        if likelihood and hasattr(self.likelihood, "set_trainable"):
            self.likelihood.set_trainable(dev)
        self._trainable = True
        return self

    def parameters(self):
        ps = list(self.kern.parameters())
        for t in (self.feature.Z, self.q_mu, self.q_sqrt):
            if isinstance(t, torch.Tensor) and t.requires_grad:
                ps.append(t)
        if hasattr(self.likelihood, "parameters"):
            ps += list(self.likelihood.parameters())
        ps += list(self.mean_function.parameters())
        return ps

    def training_loss(self, X=None, Y=None):
        """- ELBO (models.py:39-59), differentiable"""
        return -self._build_likelihood(X, Y)

    def optimize(self, iterations=100, lr=1e-2, optimizer=None, callback=None):
        """A plain Adam loop over training_loss() (minibatches as configured).  Returns the list of ELBO values."""
        if not self._trainable:
            self.set_trainable()
        opt = optimizer if optimizer is not None else torch.optim.Adam(self.parameters(), lr=lr)
        history = []
        for it in range(iterations):
            opt.zero_grad(set_to_none=True)
            loss = self.training_loss()
            loss.backward()
            opt.step()
            history.append(-float(loss.item()))
            check_cholesky()
            if callback is not None:
                callback(it, history[-1])
        self.kern.sync_trainable()
        return history

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
        f_mean, f_var, Kzz, q_mu, q_sqrt = self._build_predict(X, return_Kzz=not self.whiten, return_q=True)
        dev = f_mean.device
        KL = gauss_kl(q_mu, q_sqrt, K=Kzz)                     # q_sqrt is already lower-triangular (_build_predict)
        var_exp = self.likelihood.variational_expectations(f_mean, f_var, self._dev(Y, dev))
        scale = float(self.num_data) / float(num_samples)
        return torch.sum(var_exp) * scale - KL

    def compute_log_likelihood(self, X=None, Y=None):
        v = float(self._build_likelihood(X, Y).item())
        check_cholesky()
        return v

    def predict_f(self, X_new, full_cov=False):
        out = self._build_predict(X_new, full_cov=full_cov)
        check_cholesky()
        return out
