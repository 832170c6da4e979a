"""
Operator-level API: same names and argument lists as gpsig/signature_algs.py (:8, :37, :76, :101, :129), torch CUDA
tensors in and out, every call is one C-ABI entry point of libgpsig_b200.so.
"""
import torch

from . import _lib
from . import low_rank_calculations as _lr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(M):
    if not isinstance(M, torch.Tensor):
        M = torch.as_tensor(M)
    if not M.is_cuda:
        raise _lib.GPSigError("gpsig_b200 operators need CUDA tensors (there is no CPU path)")
    if M.dtype != torch.float32:
        M = M.to(torch.float32)
    return M


def _sigkern(M, num_levels, order, difference, upper_only=False):
    lib = _lib.load()
    M = _f32(M)
    if M.stride(-1) != 1:
        M = M.contiguous()
    if M.dim() == 4:
        n1, L1, n2, L2 = M.shape
        si, ss, sj = M.stride(0), M.stride(1), M.stride(2)
        out = torch.empty((num_levels + 1, n1, n2), device=M.device, dtype=torch.float32)
    elif M.dim() == 3:  # signature_algs.py:21-23: batch of (L1, L2) tiles
        n, L1, L2 = M.shape
        n1, n2 = 1, n
        si, ss, sj = 0, M.stride(1), M.stride(0)
        out = torch.empty((num_levels + 1, n), device=M.device, dtype=torch.float32)
    else:
        raise ValueError("M must be (n1, L1, n2, L2) or (n, L1, L2)")
    if out.numel() == 0:
        return out
    with torch.cuda.device(M.device):
        rc = lib.gpsig_sigkern_levels(M.data_ptr(), n1, L1, n2, L2, si, ss, sj, num_levels, order, int(bool(difference)),
                                      int(bool(upper_only)), out.data_ptr(), _stream())
    _lib.check(rc, "gpsig_sigkern_levels")
    return out


def signature_kern_first_order(M, num_levels, difference=True):
    """signature_algs.py:8-35.  M (n1, L1, n2, L2) or (n, L1, L2) -> (num_levels+1, n1, n2) or (num_levels+1, n)."""
    return _sigkern(M, num_levels, 1, difference)


def signature_kern_higher_order(M, num_levels, order=2, difference=True):
    """signature_algs.py:37-74."""
    order = num_levels if (order <= 0 or order >= num_levels) else order
    return _sigkern(M, num_levels, order, difference)


def tensor_kern(M, num_levels):
    """signature_algs.py:76-99.  M (T, nz, nz2) -> (num_levels+1, nz, nz2)."""
    lib = _lib.load()
    M = _f32(M).contiguous()
    T, nz, nz2 = M.shape
    if T != num_levels * (num_levels + 1) // 2:
        raise ValueError("M must have num_levels*(num_levels+1)/2 components")
    out = torch.empty((num_levels + 1, nz, nz2), device=M.device, dtype=torch.float32)
    with torch.cuda.device(M.device):
        rc = lib.gpsig_tensor_kern_levels(M.data_ptr(), num_levels, nz, nz2, 0, out.data_ptr(), _stream())
    _lib.check(rc, "gpsig_tensor_kern_levels")
    return out


def _tens_vs_seq(M, num_levels, order, difference):
    lib = _lib.load()
    M = _f32(M).contiguous()
    T, nz, n, L = M.shape
    if T != num_levels * (num_levels + 1) // 2:
        raise ValueError("M must have num_levels*(num_levels+1)/2 components")
    out = torch.empty((num_levels + 1, nz, n), device=M.device, dtype=torch.float32)
    with torch.cuda.device(M.device):
        rc = lib.gpsig_tens_vs_seq_levels(M.data_ptr(), num_levels, nz, n, L, order, int(bool(difference)), 0, out.data_ptr(),
                                          _stream())
    _lib.check(rc, "gpsig_tens_vs_seq_levels")
    return out


def signature_kern_tens_vs_seq_first_order(M, num_levels, difference=True):
    """signature_algs.py:101-127.  M (T, nz, n, L) -> (num_levels+1, nz, n)."""
    return _tens_vs_seq(M, num_levels, 1, difference)


def signature_kern_tens_vs_seq_higher_order(M, num_levels, order=2, difference=True):
    """signature_algs.py:129-160."""
    order = num_levels if (order <= 0 or order >= num_levels) else order
    return _tens_vs_seq(M, num_levels, order, difference)


def _level_seed(seeds, i):
    return None if seeds is None else seeds[i]


def signature_kern_first_order_lr_feature(U, num_levels, rank_bound, sparsity='sqrt', seeds=None, difference=True,
                                          projections=None, literal=False):
    """signature_algs.py:162-192.  U (n, L, C) low-rank features of the embedded sequences -> list of num_levels+1
    factors.  `seeds` ((num_levels-1, 2) ints) or `projections` (list of low_rank_calculations.Projection) fix the
    random projections; literal=True reproduces the reference's :191 (every level >= 2 repeats level 1)."""
    U = _f32(U)
    n = U.shape[0]
    Phi = [torch.ones((n, 1), device=U.device, dtype=torch.float32)]
    if difference:
        U = (U[:, 1:, :] - U[:, :-1, :]).contiguous()
    first = U.sum(dim=1)
    Phi.append(first)
    P = U
    for i in range(2, num_levels + 1):
        proj = projections[i - 2] if projections is not None else _lr.draw_projection(
            U.shape[-1], P.shape[-1], rank_bound, sparsity, seed=_level_seed(seeds, i - 2), device=U.device)
        P, phiP = _lr.lr_seq_level(U, P, proj)
        Phi.append(first if literal else phiP)
    return Phi


def tensor_kern_lr_feature(U, num_levels, rank_bound, sparsity='sqrt', seeds=None, projections=None):
    """signature_algs.py:194-222.  U (T, nz, C) -> list of num_levels+1 factors (nz, .)."""
    U = _f32(U)
    nz = U.shape[1]
    Phi = [torch.ones((nz, 1), device=U.device, dtype=torch.float32)]
    k = 0
    for i in range(1, num_levels + 1):
        R = U[k].contiguous()
        k += 1
        for j in range(1, i):
            proj = projections[j - 1] if projections is not None else _lr.draw_projection(
                U.shape[-1], R.shape[-1], rank_bound, sparsity, seed=_level_seed(seeds, j - 1), device=U.device)
            R = _lr.lr_hadamard_prod_rand(U[k].contiguous(), R, proj)
            k += 1
        Phi.append(R)
    return Phi
