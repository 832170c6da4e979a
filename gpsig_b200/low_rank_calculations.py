"""
Low-rank helpers -- same role and names as gpsig/low_rank_calculations.py (:12-193): Nystrom feature map and the
randomised low-rank Hadamard product (very sparse Gaussian JL projection / coordinate subsampling).

Randomness: the reference draws from TensorFlow's (stateless) random ops, which cannot be reproduced outside TF.  Here
every draw comes from a NumPy Generator on the host -- `seed` arguments play the role of the reference's `seeds` (the
same seed gives the same projection, which is what makes Phi(X) Phi(X2)^T meaningful, kernels.py:443-449) -- and every
function also accepts the draws themselves, which is how parity with the oracle / the reference's golden outputs is
tested.  The arithmetic runs on the device: projections through the C ABI (gpsig_lr_hadamard_csc, gpsig_lr_seq_level),
Grams through gpsig_gram, the C x C eigendecomposition and the feature GEMM through torch.linalg / cuBLAS.
"""
import numpy as np
import torch

from . import _lib, settings


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rng(seed):
    if isinstance(seed, np.random.Generator):
        return seed
    if seed is None:
        return np.random.default_rng()
    return np.random.default_rng(np.asarray(seed, dtype=np.uint64).reshape(-1))


def _draw_indices(n, l, rng=None):
    """low_rank_calculations.py:12-23: l indices out of n without replacement (and the rest)."""
    idx = _rng(rng).permutation(n)
    return idx[:l], idx[l:]


class Projection:
    """A random projection of the k1*k2 outer product to r components, held on the device in CSC form."""

    def __init__(self, k1, k2, r, colptr, ia, ib, val, scale, device, dense=None):
        self.k1, self.k2, self.r, self.scale = int(k1), int(k2), int(r), float(scale)
        self.colptr = torch.as_tensor(np.asarray(colptr, dtype=np.int32)).to(device)
        self.ia = torch.as_tensor(np.asarray(ia, dtype=np.int32)).to(device)
        self.ib = torch.as_tensor(np.asarray(ib, dtype=np.int32)).to(device)
        self.val = torch.as_tensor(np.asarray(val, dtype=np.float32)).to(device)
        self.dense = dense  # (k1*k2, r) float64 on the host, for tests

    @staticmethod
    def from_dense(R, k1, k2, s, device):
        """R (k1*k2, r) as drawn at low_rank_calculations.py:180; row q pairs A[q % k1] with B[q // k1] (:165-170)."""
        R = np.asarray(R, dtype=np.float64)
        D, r = R.shape
        assert D == k1 * k2
        rows, cols = np.nonzero(R)
        order = np.argsort(cols, kind="stable")
        rows, cols = rows[order], cols[order]
        colptr = np.zeros(r + 1, dtype=np.int64)
        np.add.at(colptr, cols + 1, 1)
        colptr = np.cumsum(colptr)
        return Projection(k1, k2, r, colptr, rows % k1, rows // k1, R[rows, cols], np.sqrt(float(s) / r), device, dense=R)

    @staticmethod
    def from_selection(select, signs, k1, k2, device):
        """coordinate subsampling (low_rank_calculations.py:104-127): component c = A[select[c,0]] * B[select[c,1]] * sign."""
        select = np.asarray(select, dtype=np.int64)
        r = select.shape[0]
        return Projection(k1, k2, r, np.arange(r + 1), select[:, 0], select[:, 1], np.asarray(signs, dtype=np.float64), 1.0, device)


def sparse_scale(D, sparsity):
    """low_rank_calculations.py:175-178."""
    D = float(D)
    return D / np.log(D) if sparsity == "log" else np.sqrt(D)


def draw_projection(k1, k2, rank_bound, sparsity="sqrt", seed=None, device="cuda"):
    """The draws of lr_hadamard_prod_rand (:76-90): 'lin' -> subsample + Rademacher, else sparse Gaussian with density 1/s."""
    rng = _rng(seed)
    D = k1 * k2
    if sparsity == "lin":
        comb = np.stack([np.tile(np.arange(k1), k2), np.repeat(np.arange(k2), k1)], axis=1)   # :113-117
        select = comb[rng.permutation(D)[:rank_bound]]
        signs = np.where(rng.random(rank_bound) <= 0.5, 1.0, -1.0)                               # :92-101
        return Projection.from_selection(select, signs, k1, k2, device)
    s = sparse_scale(D, sparsity)
    mask = rng.random((D, rank_bound)) <= 1.0 / s                                               # :139-149
    R = np.where(mask, rng.standard_normal((D, rank_bound)), 0.0)
    return Projection.from_dense(R, k1, k2, s, device)


def lr_hadamard_prod_rand(A, B, proj):
    """low_rank_calculations.py:76-90 with the projection given: A (..., k1), B (..., k2) -> (..., r)."""
    lib = _lib.load()
    A = A.to(torch.float32).contiguous()
    B = B.to(torch.float32).contiguous()
    assert A.shape[:-1] == B.shape[:-1] and A.shape[-1] == proj.k1 and B.shape[-1] == proj.k2
    rows = A.numel() // proj.k1
    out = torch.empty(A.shape[:-1] + (proj.r,), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        rc = lib.gpsig_lr_hadamard_csc(A.data_ptr(), rows, proj.k1, B.data_ptr(), proj.k2, proj.colptr.data_ptr(),
                                       proj.ia.data_ptr(), proj.ib.data_ptr(), proj.val.data_ptr(), proj.r, proj.scale,
                                       out.data_ptr(), _stream())
    _lib.check(rc, "gpsig_lr_hadamard_csc")
    return out


def lr_seq_level(U, P, proj):
    """signature_algs.py:182-188, one level: returns (proj(U, ExclCumsum_t(P)), its sum over time)."""
    lib = _lib.load()
    U = U.contiguous()
    P = P.contiguous()
    n, Lr, k1 = U.shape
    assert P.shape[:2] == (n, Lr) and k1 == proj.k1 and P.shape[2] == proj.k2
    out = torch.empty((n, Lr, proj.r), device=U.device, dtype=torch.float32)
    phi = torch.empty((n, proj.r), device=U.device, dtype=torch.float32)
    if n == 0:
        return out, phi
    with torch.cuda.device(U.device):
        rc = lib.gpsig_lr_seq_level(U.data_ptr(), P.data_ptr(), n, Lr, k1, proj.k2, proj.colptr.data_ptr(), proj.ia.data_ptr(),
                                    proj.ib.data_ptr(), proj.val.data_ptr(), proj.r, proj.scale, out.data_ptr(), phi.data_ptr(),
                                    _stream())
    _lib.check(rc, "gpsig_lr_seq_level")
    return out, phi


def Nystrom_map(X, kern, nys_samples=None, num_components=None, diag_draw=None, rng=None):
    """
    low_rank_calculations.py:26-61.  X (num_samples, d) device tensor of (scaled) points; kern(A, B) -> Gram on the device.
    `diag_draw` is the U[0,1) vector of :52 (quirk Q8), drawn from `rng` when absent.  The C x C eigendecomposition and
    the whitening run in float64 (C is ~50; 1 / sqrt(lambda + 1e-6) amplifies fp32 noise), the feature GEMM in fp32.
    """
    if nys_samples is None and num_components is None:
        raise ValueError('One of num_components or nys_samples should be given')
    g = _rng(rng)
    if nys_samples is None:
        idx, _ = _draw_indices(X.shape[0], num_components, g)
        nys_samples = X[torch.as_tensor(idx, device=X.device)]
    C = nys_samples.shape[0]
    if diag_draw is None:
        diag_draw = g.random(C)
    W = kern(nys_samples, nys_samples).to(torch.float64)
    W = W + torch.diag(settings.jitter * torch.as_tensor(np.asarray(diag_draw, dtype=np.float64), device=X.device))
    S, Uv = torch.linalg.eigh(W)
    Dm = torch.sqrt(S + settings.jitter)
    Wh = (Uv / Dm[None, :]).to(torch.float32)
    return kern(X, nys_samples) @ Wh
