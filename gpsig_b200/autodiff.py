"""
Differentiable route of the covariance path -- what lets `models.SVGP` train on the device.

The reference gets its gradients from TensorFlow autodiff through the whole graph (gpsig/training.py:140-203 ->
models.py:39-59 -> kernels.py -> signature_algs.py:26-33, :114-125).  Here the two recursions are custom autograd
functions whose backward passes are hand-written CUDA kernels (csrc/vjp.cu through gpsig_sigkern_levels_vjp /
gpsig_tens_vs_seq_levels_vjp: the forward state is run backwards, nothing per entry is stored), and everything around
them -- scaling by lengthscales, the static-kernel Gram (evaluated in float64, so no cancellation anywhere), the
differencing of signature_algs.py:26 / :114 / kernels.py:330, normalisation and level weights -- is ordinary tensor
algebra that torch differentiates (dense contractions on cuBLAS).

This route materialises the INCREMENT tensor of the call in float32 (the reference materialises the float64 Gram and every
R_m), so it is meant for training-sized batches (minibatches, diagonal tiles, Z x N inducing blocks); when the float64 Gram
of a call is larger than settings.autodiff_gram_budget_bytes it is evaluated in blocks under activation checkpointing.
The fused forward kernels stay the path for everything that does not need a gradient.  First order only (order == 1), exact mode only (no low-rank).
"""
import math

import numpy as np
import torch

from . import _lib, settings


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------------------------------------------------
# the two recursions as autograd functions on the INCREMENT tensors
# ----------------------------------------------------------------------------------------------------------------------
class SigKernFirstOrder(torch.autograd.Function):
    """signature_algs.py:28-33 on increments Delta (n1, L1, n2, L2) -> (num_levels + 1, n1, n2)."""

    @staticmethod
    def forward(ctx, Delta, num_levels):
        lib = _lib.load()
        Delta = Delta.contiguous()
        n1, L1, n2, L2 = Delta.shape
        out = torch.empty((num_levels + 1, n1, n2), device=Delta.device, dtype=torch.float32)
        with torch.cuda.device(Delta.device):
            rc = lib.gpsig_sigkern_levels(Delta.data_ptr(), n1, L1, n2, L2, Delta.stride(0), Delta.stride(1), Delta.stride(2),
                                          num_levels, 1, 0, 0, out.data_ptr(), _stream())
        _lib.check(rc, "gpsig_sigkern_levels")
        ctx.save_for_backward(Delta)
        ctx.num_levels = num_levels
        return out

    @staticmethod
    def backward(ctx, G):
        lib = _lib.load()
        Delta, = ctx.saved_tensors
        n1, L1, n2, L2 = Delta.shape
        G = G.contiguous().to(torch.float32)
        Dbar = torch.empty_like(Delta)
        with torch.cuda.device(Delta.device):
            rc = lib.gpsig_sigkern_levels_vjp(Delta.data_ptr(), n1, L1, n2, L2, Delta.stride(0), Delta.stride(1),
                                              Delta.stride(2), ctx.num_levels, G.data_ptr(), Dbar.data_ptr(), _stream())
        _lib.check(rc, "gpsig_sigkern_levels_vjp")
        return Dbar, None


class TensVsSeqFirstOrder(torch.autograd.Function):
    """signature_algs.py:116-125 on increments H (T, nz, n, Lh) -> (num_levels + 1, nz, n)."""

    @staticmethod
    def forward(ctx, H, num_levels):
        lib = _lib.load()
        H = H.contiguous()
        T, nz, n, Lh = H.shape
        out = torch.empty((num_levels + 1, nz, n), device=H.device, dtype=torch.float32)
        with torch.cuda.device(H.device):
            rc = lib.gpsig_tens_vs_seq_levels(H.data_ptr(), num_levels, nz, n, Lh, 1, 0, 0, out.data_ptr(), _stream())
        _lib.check(rc, "gpsig_tens_vs_seq_levels")
        ctx.save_for_backward(H)
        ctx.num_levels = num_levels
        return out

    @staticmethod
    def backward(ctx, G):
        lib = _lib.load()
        H, = ctx.saved_tensors
        T, nz, n, Lh = H.shape
        G = G.contiguous().to(torch.float32)
        Hbar = torch.empty_like(H)
        with torch.cuda.device(H.device):
            rc = lib.gpsig_tens_vs_seq_levels_vjp(H.data_ptr(), ctx.num_levels, nz, n, Lh, G.data_ptr(), Hbar.data_ptr(), _stream())
        _lib.check(rc, "gpsig_tens_vs_seq_levels_vjp")
        return Hbar, None


def sigkern_first_order(M, num_levels, difference=True):
    """signature_algs.py:8-35, differentiable.  M (n1, L1, n2, L2) or (n, L1, L2)."""
    M = M.to(torch.float32)
    three = M.dim() == 3
    if three:
        M = M[:, :, None, :]                                   # (n, L1, 1, L2): pairs (i, 0)
    if difference:                                             # signature_algs.py:26
        M = M[:, 1:, :, 1:] + M[:, :-1, :, :-1] - M[:, :-1, :, 1:] - M[:, 1:, :, :-1]
    n1, L1, n2, L2 = M.shape
    if L1 < 1 or L2 < 1:
        out = torch.zeros((num_levels + 1, n1, n2), device=M.device, dtype=torch.float32)
        out[0] = 1.0
    else:
        out = SigKernFirstOrder.apply(M, num_levels)
    return out[:, :, 0] if three else out


def tens_vs_seq_first_order(M, num_levels, difference=True):
    """signature_algs.py:101-127, differentiable.  M (T, nz, n, L)."""
    M = M.to(torch.float32)
    if difference:                                             # signature_algs.py:114
        M = M[..., 1:] - M[..., :-1]
    if M.shape[-1] < 1:
        out = torch.zeros((num_levels + 1,) + tuple(M.shape[1:3]), device=M.device, dtype=torch.float32)
        out[0] = 1.0
        return out
    return TensVsSeqFirstOrder.apply(M, num_levels)


def tensor_kern(M, num_levels):
    """signature_algs.py:76-99 in tensor algebra (the Z x Z side is tiny).  M (T, nz, nz2)."""
    levels = [torch.ones_like(M[0])]
    k = 0
    for m in range(1, num_levels + 1):
        levels.append(torch.prod(M[k:k + m], dim=0))
        k += m
    return torch.stack(levels, dim=0)


# ----------------------------------------------------------------------------------------------------------------------
# static kernels (kernels.py:786-993) on scaled points, float64
# ----------------------------------------------------------------------------------------------------------------------
def _sqdist(A, B):
    d = (A * A).sum(-1)[..., :, None] + (B * B).sum(-1)[..., None, :] - 2.0 * A @ B.transpose(-1, -2)   # kernels.py:765-776
    return torch.clamp(d, min=0.0)


def base_gram(kern, A, B=None):
    """Static-kernel Gram of scaled points A (..., r1, d), B (..., r2, d) in float64 (kernels.py:225-230)."""
    A = A.to(torch.float64)
    B = A if B is None else B.to(torch.float64)
    kind = kern._kind
    if kind == "linear":
        return A @ B.transpose(-1, -2)
    if kind == "rbf":
        return torch.exp(-0.5 * _sqdist(A, B))
    if kind == "cosine":
        na, nb = torch.sqrt((A * A).sum(-1)), torch.sqrt((B * B).sum(-1))
        return (A @ B.transpose(-1, -2)) / (na[..., :, None] * nb[..., None, :])
    if kind == "poly":
        return (A @ B.transpose(-1, -2) + kern._tparam("gamma_poly", A.device)) ** kern._tparam("degree", A.device)
    if kind == "mix":
        mix = kern._tparam("mixing", A.device)
        return mix * torch.exp(-0.5 * _sqdist(A, B)) + (1.0 - mix) * (A @ B.transpose(-1, -2))
    if kind in ("matern12", "matern32", "matern52"):
        r = torch.sqrt(torch.clamp(_sqdist(A, B), min=1e-40))                                          # kernels.py:779-781
        if kind == "matern12":
            return torch.exp(-r)
        if kind == "matern32":
            return (1.0 + math.sqrt(3.0) * r) * torch.exp(-math.sqrt(3.0) * r)
        return (1.0 + math.sqrt(5.0) * r + 5.0 / 3.0 * r * r) * torch.exp(-math.sqrt(5.0) * r)
    raise NotImplementedError("no differentiable route for the %s static kernel" % kind)


# ----------------------------------------------------------------------------------------------------------------------
# kernels.py:188-340, :430-476 -- the pieces SignatureKernel composes, differentiable
# ----------------------------------------------------------------------------------------------------------------------
def scale(kern, X, tensors=False):
    """kernels.py:357-361 (sequences) / :366-398 (inducing tensors): X / lengthscales on the last axis, x gamma for lagged
    copies (tensors: only when there are lengthscales, as in the reference)."""
    inv = kern._inv_ls_tensor(X.device, tensors)
    return X.to(torch.float64) if inv is None else X.to(torch.float64) * inv


def add_lags(kern, X):
    """lags.py:7-63: lagged copies by linear interpolation as extra features.  X (n, L, d) -> (n, L, (P + 1) d)."""
    if kern.num_lags == 0:
        return X
    n, L, d = X.shape
    lags = kern._tparam("lags", X.device)
    time = torch.arange(L, device=X.device, dtype=torch.float64) / float(L - 1)
    tq = torch.clamp(time[:, None] - lags[None, :], min=0.0)                                       # (L, P)
    left = torch.clamp(torch.floor((tq + kern.jitter) * (L - 1) + 1e-9).to(torch.int64), 0, L - 2).detach()
    tl, tr = time[left], time[left + 1]
    Xd = X.to(torch.float64)
    Xl, Xr = Xd[:, left, :], Xd[:, left + 1, :]                                                    # (n, L, P, d)
    Xq = Xl + ((tq - tl) / (tr - tl))[None, :, :, None] * (Xr - Xl)
    return torch.cat((Xd[:, :, None, :], Xq), dim=2).reshape(n, L, -1)


def _maybe_checkpoint(fn, *tensors):
    """Run fn(*tensors) so that autograd keeps only its INPUTS: the float64 Gram of a block and the three or four
    same-sized intermediates torch would save for its backward (squared distances, the clamp's input, the exponential's
    output, the differences) are recomputed when the backward pass reaches the block instead of staying resident --
    the Gram blocks are what limits the batch size of a training step, not their arithmetic."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        from torch.utils.checkpoint import checkpoint
        return checkpoint(fn, *tensors, use_reentrant=False)
    return fn(*tensors)


def K_seq_levels(kern, X, X2=None, rows_per_chunk=None):
    """kernels.py:208-237 on RAW sequences (n, L, d); returns (M + 1, n1, n2)."""
    Xs = scale(kern, X)
    X2s = Xs if X2 is None else scale(kern, X2)
    n1, L1, d = Xs.shape
    n2, L2 = X2s.shape[0], X2s.shape[1]
    flat2 = X2s.reshape(n2 * L2, d)
    big = 8 * n1 * L1 * n2 * L2 > settings.autodiff_gram_budget_bytes                               # float64 Gram of the call
    if rows_per_chunk is None:
        rows_per_chunk = max(1, int((1 << 27) // max(1, L1 * n2 * L2))) if big else n1              # blocks of <= 1 GB
    difference = kern.difference

    def block(Xc, flat):                                                                          # float32 increments of a row block
        M = base_gram(kern, Xc.reshape(Xc.shape[0] * L1, d), flat).reshape(Xc.shape[0], L1, n2, L2)
        if difference:                                                                             # signature_algs.py:26
            M = M[:, 1:, :, 1:] + M[:, :-1, :, :-1] - M[:, :-1, :, 1:] - M[:, 1:, :, :-1]
        return M.to(torch.float32)

    outs = []
    for c0 in range(0, n1, rows_per_chunk):
        c1 = min(n1, c0 + rows_per_chunk)
        M = _maybe_checkpoint(block, Xs[c0:c1], flat2) if big else block(Xs[c0:c1], flat2)
        outs.append(sigkern_first_order(M, kern.num_levels, difference=False))
    return torch.cat(outs, dim=1)


def K_seq_diag_levels(kern, X):
    """kernels.py:188-205; returns (M + 1, n)."""
    Xs = scale(kern, X)
    M = base_gram(kern, Xs, Xs)                                                                    # (n, L, L) batched
    return sigkern_first_order(M, kern.num_levels, kern.difference)


def K_tens_levels(kern, Z, increments=False):
    """kernels.py:263-283 on RAW tensors (T, nz, [2,] d); returns (M + 1, nz, nz)."""
    Zs = scale(kern, Z, tensors=True)
    T, nz, d = Zs.shape[0], Zs.shape[1], Zs.shape[-1]
    if increments:
        M = base_gram(kern, Zs.reshape(T, 2 * nz, d)).reshape(T, nz, 2, nz, 2)
        M = M[:, :, 1, :, 1] + M[:, :, 0, :, 0] - M[:, :, 1, :, 0] - M[:, :, 0, :, 1]               # kernels.py:276-277
    else:
        M = base_gram(kern, Zs)
    return tensor_kern(M.to(torch.float32), kern.num_levels)


def K_tens_vs_seq_levels(kern, Z, X, increments=False):
    """kernels.py:313-340 on RAW tensors / sequences; returns (M + 1, nz, n)."""
    Zs, Xs = scale(kern, Z, tensors=True), scale(kern, X)
    T, nz, d = Zs.shape[0], Zs.shape[1], Zs.shape[-1]
    n, L = Xs.shape[0], Xs.shape[1]
    Zflat = Zs.reshape(-1, d)
    difference = kern.difference

    def block(Zf, Xc):                                                                            # float32 increments of a block of sequences
        nc = Xc.shape[0]
        M = base_gram(kern, Zf, Xc.reshape(nc * L, d))
        if increments:
            M = M.reshape(T, nz, 2, nc, L)
            M = M[:, :, 1] - M[:, :, 0]                                                            # kernels.py:330
        else:
            M = M.reshape(T, nz, nc, L)
        if difference:                                                                             # signature_algs.py:114
            M = M[..., 1:] - M[..., :-1]
        return M.to(torch.float32)

    per_seq = max(1, Zflat.shape[0] * L)                                                           # float64 Gram entries per sequence
    if 8 * per_seq * n > settings.autodiff_gram_budget_bytes:                                      # blocks of <= 512 MB, recomputed in backward
        chunk = max(1, int((1 << 26) // per_seq))
        parts = [_maybe_checkpoint(block, Zflat, Xs[c0:min(n, c0 + chunk)]) for c0 in range(0, n, chunk)]
        M = parts[0] if len(parts) == 1 else torch.cat(parts, dim=2)
    else:
        M = block(Zflat, Xs)
    return tens_vs_seq_first_order(M, kern.num_levels, difference=False)


def finish(kern, levels, diag1=None, diag2=None, symmetric=False, normalize=True, return_levels=False, unit_weights=False):
    """kernels.py:430-433 / :455-469 / :471-476."""
    lv = levels
    if normalize:
        if symmetric:                                                                              # kernels.py:431-433
            n = lv.shape[1]
            lv = lv + kern.jitter * torch.eye(n, device=lv.device, dtype=lv.dtype)[None]
            dsq = torch.sqrt(torch.diagonal(lv, dim1=1, dim2=2))
            lv = lv / (dsq[:, :, None] * dsq[:, None, :])
        else:                                                                                      # kernels.py:463-469, :578-581
            if diag1 is not None:
                lv = lv / torch.sqrt(diag1 + kern.jitter)[:, :, None]
            if diag2 is not None:
                lv = lv / torch.sqrt(diag2 + kern.jitter)[:, None, :]
    if not unit_weights:
        w = kern._weights_tensor(lv.device).to(lv.dtype)
        lv = lv * w.reshape((-1,) + (1,) * (lv.dim() - 1))
    return lv if return_levels else lv.sum(dim=0)


def inv_softplus(y):
    """inverse of gpflow's transforms.positive (softplus with lower = 1e-6)."""
    y = np.asarray(y, dtype=np.float64) - 1e-6
    return np.where(y > 30.0, y, np.log(np.expm1(np.maximum(y, 1e-300))))


def softplus(raw):
    return torch.nn.functional.softplus(raw) + 1e-6
