"""
Inducing variables and the Kuu / Kuf / Kuu_Kuf_Kff evaluations -- same classes and function signatures as
gpsig/inducing_variables.py (:14-137).  The reference registers these with gpflow's multiple dispatch keyed on
(InducingTensors | InducingSequences, SignatureKernel[, object]); here the dispatch is an isinstance check.
"""
import numpy as np
import torch

from .kernels import SignatureKernel


class SignatureInducing:
    """inducing_variables.py:14-26."""

    def __init__(self, Z, num_levels, learn_weights=False):
        self.Z = Z if isinstance(Z, torch.Tensor) else np.asarray(Z, dtype=np.float64)
        self.learn_weights = learn_weights
        if learn_weights:
            self.W = np.tile(np.eye(self.__len__())[None, ...], [num_levels, 1, 1])

    def __len__(self):
        return self.Z.shape[0]


class InducingTensors(SignatureInducing):
    """inducing_variables.py:28-49.  Z: (T, num_tensors, d) or, with increments, (T, num_tensors, 2, d)."""

    def __init__(self, Z, num_levels, increments=False, **kwargs):
        len_tensors = int(num_levels * (num_levels + 1) / 2)
        assert Z.shape[0] == len_tensors
        if increments:
            assert Z.ndim == 4
            assert Z.shape[2] == 2
        super().__init__(Z, num_levels, **kwargs)
        self.len_tensors = len_tensors
        self.increments = increments

    def __len__(self):
        return self.Z.shape[1]


class InducingSequences(SignatureInducing):
    """inducing_variables.py:89-98.  Z: (num_inducing, len_inducing, num_features)."""

    def __init__(self, Z, num_levels, **kwargs):
        super().__init__(Z, num_levels, **kwargs)
        self.len_inducing = Z.shape[1]


def _W(feat, dev, dtype=torch.float32):
    if isinstance(feat.W, torch.Tensor):
        return feat.W.to(device=dev, dtype=dtype)
    return torch.as_tensor(np.asarray(feat.W, dtype=np.float64)).to(device=dev, dtype=dtype)


def _eye(n, dev, dtype=torch.float32):
    return torch.eye(n, device=dev, dtype=dtype)


def _check(feat, kern):
    if not isinstance(kern, SignatureKernel):
        raise NotImplementedError("Kuu/Kuf are defined for SignatureKernel only")
    if not isinstance(feat, (InducingTensors, InducingSequences)):
        raise NotImplementedError("feat must be InducingTensors or InducingSequences")


def Kuu_Kuf_Kff(feat, kern, X_new, *, jitter=0.0, full_f_cov=False, literal=False):
    """inducing_variables.py:51-66 (tensors) / :122-137 (sequences).  full_f_cov=True raises NameError in the reference
    (tf.shape(X) with X undefined, quirk Q5); the evident intent (jitter * I on Kxx) is implemented.  literal=True
    (sequences only) reproduces the reference's double normalisation of Kzx (kernels.py:713 then :750, quirk Q4) -- what
    the golden vectors of the unmodified reference hold; the default divides once."""
    _check(feat, kern)
    seq = isinstance(feat, InducingSequences)
    lv = bool(feat.learn_weights)
    if seq:
        Z = feat.Z.reshape(len(feat), -1)
        Kzz, Kzx, Kxx = kern.K_seq_n_seq_covs(Z, X_new, full_X2_cov=full_f_cov, return_levels=lv, literal=literal)
    else:
        Kzz, Kzx, Kxx = kern.K_tens_n_seq_covs(feat.Z, X_new, full_X_cov=full_f_cov, return_levels=lv,
                                               increments=feat.increments)
    if lv:
        W = _W(feat, Kzz.device, Kzz.dtype)
        Kzz = Kzz[0] + torch.sum(torch.matmul(torch.matmul(W, Kzz[1:]), W.transpose(-1, -2)), dim=0)
        Kzx = Kzx[0] + torch.sum(torch.matmul(W, Kzx[1:]), dim=0)
        Kxx = torch.sum(Kxx, dim=0)
    Kzz = Kzz + jitter * _eye(len(feat), Kzz.device, Kzz.dtype)
    if full_f_cov:
        Kxx = Kxx + jitter * _eye(Kxx.shape[-1], Kxx.device, Kxx.dtype)
    else:
        Kxx = Kxx + jitter
    return Kzz, Kzx, Kxx


def Kuf(feat, kern, X_new):
    """inducing_variables.py:68-76 / :112-120."""
    _check(feat, kern)
    lv = bool(feat.learn_weights)
    if isinstance(feat, InducingSequences):
        Z = feat.Z.reshape(len(feat), -1)
        Kzx = kern.K(Z, X_new, presliced_X=True, return_levels=lv)
    else:
        Kzx = kern.K_tens_vs_seq(feat.Z, X_new, return_levels=lv, increments=feat.increments)
    if lv:
        Kzx = Kzx[0] + torch.sum(torch.matmul(_W(feat, Kzx.device, Kzx.dtype), Kzx[1:]), dim=0)
    return Kzx


def Kuu(feat, kern, *, jitter=0.0, full_f_cov=False):
    """inducing_variables.py:78-87 / :101-110."""
    _check(feat, kern)
    lv = bool(feat.learn_weights)
    if isinstance(feat, InducingSequences):
        Z = feat.Z.reshape(len(feat), -1)
        Kzz = kern.K(Z, presliced=True, return_levels=lv)
    else:
        Kzz = kern.K_tens(feat.Z, return_levels=lv, increments=feat.increments)
    if lv:
        W = _W(feat, Kzz.device, Kzz.dtype)
        Kzz = Kzz[0] + torch.sum(torch.matmul(torch.matmul(W, Kzz[1:]), W.transpose(-1, -2)), dim=0)
    return Kzz + jitter * _eye(len(feat), Kzz.device, Kzz.dtype)
