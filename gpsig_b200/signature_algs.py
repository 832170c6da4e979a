"""
Operator-level API: same names and argument lists as gpsig/signature_algs.py (:8, :37, :76, :101, :129), torch CUDA
tensors in and out, every call is one C-ABI entry point of libgpsig_b200.so.
"""
import torch

from . import _lib


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
