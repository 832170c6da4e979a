"""
SignatureKernel family -- same constructor kwargs (kernels.py:18-19), same 2-D (N, L*d) input convention
(kernels.py:417-419), same public methods (K :401, Kdiag :479, K_tens :513, K_tens_vs_seq :539, K_tens_n_seq_covs :591,
K_seq_n_seq_covs :674) and numpy-facing compute_* helpers (:141-186) as the reference.  Every method is a short
sequence of C-ABI calls into libgpsig_b200.so; torch tensors only hold device memory.

Parameters (variances, sigma, lengthscales, ...) are plain numpy values in constrained space, i.e. what the reference
reads inside @params_as_tensors.
"""
import numpy as np
import torch

from . import _lib, settings
from . import autodiff as _ad
from . import low_rank_calculations as _lr
from . import signature_algs as _algs

_MAX_FUSED_FEATURES = 16   # widest (lagged) state space of the fused Gram producers (gram.cu)
_KIND = dict(linear=0, rbf=1, cosine=2, poly=3, mix=4, matern12=5, matern32=6, matern52=7, spectral=8)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()



_CONST_CACHE = {}


def _dev_const(arr, dev):
    """Small read-only parameter vectors (lengthscale multipliers, level weights) on the device, cached by VALUE: a fresh
    pageable host-to-device copy per call is a driver round trip on the critical path of every step."""
    key = (str(dev), arr.dtype.str, arr.shape, arr.tobytes())
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) > 256:
            _CONST_CACHE.clear()
        t = torch.as_tensor(arr).to(dev)
        _CONST_CACHE[key] = t
    return t

class SignatureKernel:
    """Base class (kernels.py:15).  Subclasses set `_kind` and static-kernel parameters."""

    _kind = None

    def __init__(self, input_dim, num_features, num_levels, active_dims=None, variances=1, lengthscales=1, order=1,
                 normalization=True, difference=True, num_lags=None, low_rank=False, num_components=50, rank_bound=None,
                 sparsity='sqrt', name=None, device=None):
        self.input_dim = int(input_dim)
        self.active_dims = None if active_dims is None else np.asarray(active_dims, dtype=np.int64)
        self.name = name
        self.num_features = int(num_features)
        self.num_levels = int(num_levels)
        self.len_examples = self._validate_number_of_features(input_dim, num_features)
        self.order = num_levels if (order <= 0 or order >= num_levels) else int(order)          # kernels.py:57
        if self.order != 1 and low_rank:                                                         # kernels.py:59-60
            raise NotImplementedError('Higher-order algorithms not compatible with low-rank mode (yet).')
        self.normalization = bool(normalization)
        self.difference = bool(difference)
        self.variances = self._validate_signature_param("variances", variances, num_levels + 1)
        self.sigma = 1.0
        self.low_rank, self.num_components, self.rank_bound, self.sparsity = self._validate_low_rank_params(
            low_rank, num_components, rank_bound, sparsity)
        if num_lags is None:
            self.num_lags = 0
        else:
            if not isinstance(num_lags, int) or num_lags < 0:                                    # kernels.py:74-75
                raise ValueError('The variable num_lags most be a nonnegative integer or None.')
            self.num_lags = int(num_lags)
            if num_lags > 0:
                self.lags = 0.1 * np.asarray(range(1, num_lags + 1), dtype=np.float64)
                gamma = 1. / np.asarray(range(1, self.num_lags + 2), dtype=np.float64)
                self.gamma = gamma / np.sum(gamma)
        if lengthscales is not None:
            self.lengthscales = self._validate_signature_param("lengthscales", lengthscales, self.num_features)
        else:
            self.lengthscales = None
        self.jitter = settings.jitter
        self.device = torch.device(device) if device is not None else None
        self._ws = None
        self.lr_rng = np.random.default_rng()   # source of the low-rank mode's draws (reseed for reproducibility)
        self._raw = {}                           # trainable parameters in unconstrained space (set_trainable)

    # ---- validators (kernels.py:94-133) ----
    def _validate_number_of_features(self, input_dim, num_features):
        if input_dim % num_features == 0:
            return int(input_dim / num_features)
        raise ValueError("The arguments num_features and input_dim are not consistent.")

    def _validate_low_rank_params(self, low_rank, num_components, rank_bound, sparsity):
        if low_rank is not None and low_rank is True:
            if sparsity not in ['log', 'sqrt', 'lin']:
                raise ValueError("Unknown sparsity argument %s. Possible values are 'sqrt', 'log', 'lin'" % sparsity)
            if rank_bound is not None and rank_bound <= 0:
                raise ValueError("The rank-bound in the low-rank algorithm must be either None or a positiv integer.")
            if num_components is None or num_components <= 0:
                raise ValueError("The number of components in the kernel approximation must be a positive integer.")
            if rank_bound is None:
                rank_bound = num_components
        elif low_rank is not None and low_rank is not False and low_rank:
            raise ValueError("Unknown low-rank argument: %s. It should be True of False." % low_rank)
        else:
            low_rank = False
        return low_rank, num_components, rank_bound, sparsity

    def _validate_signature_param(self, name, value, length):
        value = value * np.ones(length, dtype=np.float64)
        correct_shape = () if length == 1 else (length,)
        if np.asarray(value).squeeze().shape != correct_shape:
            raise ValueError("shape of parameter {} is not what is expected ({})".format(name, length))
        return value

    # ---- host helpers ----
    def _dev(self, like=None):
        if isinstance(like, torch.Tensor) and like.is_cuda:
            return like.device
        if self.device is not None:
            return self.device
        if not torch.cuda.is_available():
            raise _lib.GPSigError("gpsig_b200 needs a CUDA device: there is no CPU implementation of the covariance path")
        return torch.device("cuda", torch.cuda.current_device())

    def _to_dev(self, X, like=None):
        dev = self._dev(X if like is None else like)
        if isinstance(X, torch.Tensor):
            return X.to(device=dev, dtype=torch.float32)
        return torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).to(dev)

    def _slice(self, X):
        """gpflow Kernel._slice (kernels.py:411): pick active_dims out of the last axis."""
        if self.active_dims is None:
            return X
        idx = torch.as_tensor(self.active_dims, device=X.device)
        return X.index_select(-1, idx)

    def _seqs(self, X, presliced=False):
        """(N, L*d) -> (N, L, d_eff): reshape (kernels.py:417-419) and, with lags, the lagged copies as extra features
        (kernels.py:352-353 -> lags.py:41-63); scaling by lengthscales / gamma happens inside the device pipeline."""
        X = self._to_dev(X)
        if not presliced:
            X = self._slice(X)
        X = X.reshape(X.shape[0], -1, self.num_features).contiguous()
        if self.num_lags > 0:
            lib = _lib.load()
            n, L, d = X.shape
            lags = torch.as_tensor(np.asarray(self.lags, dtype=np.float32)).to(X.device)
            out = torch.empty((n, L, (self.num_lags + 1) * d), device=X.device, dtype=torch.float32)
            if n > 0:
                with torch.cuda.device(X.device):
                    rc = lib.gpsig_add_lags(X.data_ptr(), n, L, d, lags.data_ptr(), self.num_lags, out.data_ptr(), _stream())
                _lib.check(rc, "gpsig_add_lags")
            X = out
        return X

    def _inv_ls(self, dev, tensors=False):
        """per-feature multiplier of the (lagged) state space: gamma[p] / lengthscales[c]  (kernels.py:357-361).  Inducing
        tensors (`tensors`) are only touched when there are lengthscales -- the lag weights included (kernels.py:374-379,
        :391-395), unlike sequences, which always get the lag weights (:360-361)."""
        if self.lengthscales is None and (self.num_lags == 0 or tensors):
            return None
        inv = np.ones(self.num_features) if self.lengthscales is None else 1.0 / np.asarray(self.lengthscales, dtype=np.float64)
        if self.num_lags > 0:
            inv = (np.asarray(self.gamma, dtype=np.float64)[:, None] * inv[None, :]).reshape(-1)
        return _dev_const(inv.astype(np.float32), dev)

    # ---- trainable parameters (the reference: gpflow Parameters with transforms.positive, kernels.py:65-66, :86) ----
    _POSITIVE = ("variances", "sigma", "lengthscales")

    def set_trainable(self, names=("variances", "sigma", "lengthscales"), device=None):
        """Turn the named parameters into torch leaf tensors in UNCONSTRAINED space (softplus transform like gpflow's
        transforms.positive; `lags` uses the logistic transform onto (0, 0.5) of kernels.py:80).  From then on every
        covariance method called with autograd enabled goes through the differentiable route (autodiff.py)."""
        dev = torch.device(device) if device is not None else self._dev()
        for nm in names:
            if nm == "lengthscales" and self.lengthscales is None:
                continue
            if nm in ("lags", "gamma") and self.num_lags == 0:
                continue
            val = np.asarray(getattr(self, nm), dtype=np.float64)
            if nm == "lags":
                raw = np.log(val / 0.5) - np.log1p(-val / 0.5)
            elif nm in self._POSITIVE or nm == "gamma":
                raw = _ad.inv_softplus(val)
            else:
                raw = val
            self._raw[nm] = torch.tensor(raw, dtype=torch.float64, device=dev, requires_grad=True)
        return self

    def parameters(self):
        return list(self._raw.values())

    def _tparam(self, nm, dev):
        """constrained value of a parameter as a float64 tensor on `dev` (a function of the raw leaf when trainable)"""
        attr = {"gamma_poly": "gamma_poly"}.get(nm, nm)
        if attr in self._raw:
            raw = self._raw[attr].to(dev)
            if attr == "lags":
                return 0.5 * torch.sigmoid(raw)
            return _ad.softplus(raw) if (attr in self._POSITIVE or attr == "gamma") else raw
        return torch.as_tensor(np.asarray(getattr(self, attr), dtype=np.float64), device=dev)

    def sync_trainable(self):
        """copy the current constrained values of the trainable parameters into the plain (numpy) attributes the
        non-differentiable fast path reads"""
        for nm in self._raw:
            with torch.no_grad():
                v = self._tparam(nm, self._raw[nm].device).cpu().numpy()
            setattr(self, nm, float(v) if v.ndim == 0 else v)

    def _grad_mode(self, *inputs):
        if not torch.is_grad_enabled():
            return False
        return bool(self._raw) or any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs)

    def _check_grad_supported(self):
        if self.order != 1:
            raise NotImplementedError("the differentiable route covers the first-order recursions (order == 1)")
        if self.low_rank:
            raise NotImplementedError("the differentiable route covers the exact mode (low_rank == False)")

    def _inv_ls_tensor(self, dev, tensors=False):
        if self.lengthscales is None and (self.num_lags == 0 or tensors):
            return None
        inv = torch.ones(self.num_features, device=dev, dtype=torch.float64) if self.lengthscales is None \
            else 1.0 / self._tparam("lengthscales", dev).reshape(-1)
        if self.num_lags > 0:
            inv = (self._tparam("gamma", dev)[:, None] * inv[None, :]).reshape(-1)
        return inv

    def _weights_tensor(self, dev):
        return self._tparam("sigma", dev) * self._tparam("variances", dev)                       # kernels.py:471

    def _to_dev64(self, X, like=None):
        """inputs of the differentiable route: float64 on the device (numpy is never rounded to fp32 first)"""
        dev = self._dev(X if like is None else like)
        if isinstance(X, torch.Tensor):
            return X.to(device=dev, dtype=torch.float64)
        return torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)).to(dev)

    def _seqs_t(self, X, presliced=False):
        """differentiable _seqs: slice, reshape, lagged copies (tensor algebra)"""
        X = self._to_dev64(X)
        if not presliced:
            X = self._slice(X)
        X = X.reshape(X.shape[0], -1, self.num_features)
        return _ad.add_lags(self, X)

    def _tens_t(self, Z, like=None):
        return self._to_dev64(Z, like=like)

    def _weights(self, dev):
        w = float(self.sigma) * np.asarray(self.variances, dtype=np.float64)                    # kernels.py:471
        return _dev_const(w.astype(np.float32), dev)

    def _static_params(self):
        return None

    def _params_ptr(self):
        p = self._static_params()
        if p is None:
            return None, 0
        import ctypes
        arr = (ctypes.c_float * len(p))(*[float(v) for v in p])
        return arr, ctypes.cast(arr, ctypes.c_void_p).value

    def _check_supported(self):
        if self._kind is None:
            raise NotImplementedError("use a SignatureKernel subclass (SignatureLinear, SignatureRBF, ...)")

    def _workspace(self, dev, n1, L1, n2, L2, d):
        lib = _lib.load()
        need = lib.gpsig_seq_kern_workspace_bytes(n1, L1, n2, L2, d, int(settings.workspace_budget_bytes))
        if self._ws is not None and self._ws.device == dev and self._ws.numel() >= need:
            return self._ws        # steady state: no driver query on the step's critical path
        free, _ = torch.cuda.mem_get_info(dev)
        budget = int(min(settings.workspace_budget_bytes, max(free // 2, 64 << 20)))
        need = lib.gpsig_seq_kern_workspace_bytes(n1, L1, n2, L2, d, budget)
        if self._ws is None or self._ws.device != dev or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    def _workspace_diag(self, dev, n, L, d):
        lib = _lib.load()
        need = lib.gpsig_seq_kern_diag_workspace_bytes(n, L, d, 256 << 20)
        if self._ws is None or self._ws.device != dev or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    # ---- device pieces (each = one C-ABI call) ----
    def _K_seq(self, X, X2=None, row_blocks=None):
        """kernels.py:208-237 on RAW (unscaled) sequences (N, L, d); returns level stack (M+1, N, N2).

        row_blocks (symmetric only): list of (begin, end) global row ranges owned by this GPU; the result is then the
        compact (M+1, sum of block sizes, N) stack holding only the entries j >= i of those rows (parallel.py)."""
        lib = _lib.load()
        dev = X.device
        n1, L1, d = X.shape
        if d > _MAX_FUSED_FEATURES:
            return self._K_seq_wide(X, X2, row_blocks)
        n2, L2 = (n1, L1) if X2 is None else (X2.shape[0], X2.shape[1])
        blocks = [(0, n1)] if row_blocks is None else list(row_blocks)
        nrows = sum(e - b for b, e in blocks)
        out = torch.empty((self.num_levels + 1, nrows, n2), device=dev, dtype=torch.float32)
        ws = self._workspace(dev, n1, L1, n2, L2, d)
        inv_ls = self._inv_ls(dev)
        keep, pptr = self._params_ptr()
        with torch.cuda.device(dev):
            if row_blocks is None:
                rc = lib.gpsig_seq_kern_levels(_KIND[self._kind], pptr, X.data_ptr(), n1, L1, _ptr(X2), n2, L2, d, _ptr(inv_ls),
                                               self.num_levels, self.order, int(self.difference), 0, n1, out.data_ptr(), 0,
                                               nrows, int(X2 is None), ws.data_ptr(), ws.numel(), _stream())
                _lib.check(rc, "gpsig_seq_kern_levels")
            else:
                import ctypes
                flat = (ctypes.c_int * (2 * len(blocks)))(*[int(v) for be in blocks for v in be])
                rc = lib.gpsig_seq_kern_levels_blocks(_KIND[self._kind], pptr, X.data_ptr(), n1, L1, _ptr(X2), n2, L2, d,
                                                      _ptr(inv_ls), self.num_levels, self.order, int(self.difference), flat,
                                                      len(blocks), out.data_ptr(), nrows, ws.data_ptr(), ws.numel(), _stream())
                _lib.check(rc, "gpsig_seq_kern_levels_blocks")
        return out

    def _K_seq_wide(self, X, X2=None, row_blocks=None):
        """_K_seq for state spaces wider than the fused producers handle (d > 16, e.g. RNN features): the static-kernel
        Gram of a row block is materialised exactly as kernels.py:225-230 does (gpsig_gram) and handed to the
        operator-level recursion (gpsig_sigkern_levels, differencing fused).  Functional fallback, not the tuned path."""
        n1, L1, d = X.shape
        Xs = self._scale_tens(X)
        X2s = Xs if X2 is None else self._scale_tens(X2)
        n2, L2 = X2s.shape[0], X2s.shape[1]
        flat2 = X2s.reshape(n2 * L2, d)
        blocks = [(0, n1)] if row_blocks is None else list(row_blocks)
        rows_per = max(1, int((1 << 30) // max(1, 4 * L1 * n2 * L2)))
        outs = []
        for b, e in blocks:
            for c0 in range(b, e, rows_per):
                c1 = min(e, c0 + rows_per)
                M = self._base_gram(Xs[c0:c1].reshape((c1 - c0) * L1, d), flat2).reshape(c1 - c0, L1, n2, L2)
                outs.append(_algs._sigkern(M, self.num_levels, self.order, self.difference))
        return torch.cat(outs, dim=1).contiguous()

    def _K_seq_diag_wide(self, X):
        n, L, d = X.shape
        Xs = self._scale_tens(X)
        outs = []
        for c0 in range(0, n, 8):
            c = min(8, n - c0)
            M = self._base_gram(Xs[c0:c0 + c].reshape(c * L, d)).reshape(c, L, c, L)
            idx = torch.arange(c, device=X.device)
            outs.append(_algs._sigkern(M[idx, :, idx, :].contiguous(), self.num_levels, self.order, self.difference))
        return torch.cat(outs, dim=1).contiguous()

    def _K_seq_diag(self, X):
        """kernels.py:188-205; returns (M+1, N)."""
        lib = _lib.load()
        dev = X.device
        n, L, d = X.shape
        if d > _MAX_FUSED_FEATURES:
            return self._K_seq_diag_wide(X)
        out = torch.empty((self.num_levels + 1, n), device=dev, dtype=torch.float32)
        ws = self._workspace_diag(dev, n, L, d)
        inv_ls = self._inv_ls(dev)
        keep, pptr = self._params_ptr()
        with torch.cuda.device(dev):
            rc = lib.gpsig_seq_kern_diag_levels(_KIND[self._kind], pptr, X.data_ptr(), n, L, d, _ptr(inv_ls), self.num_levels,
                                                self.order, int(self.difference), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _stream())
        _lib.check(rc, "gpsig_seq_kern_diag_levels")
        return out

    def _finish(self, levels, diag1=None, diag2=None, symmetric=False, normalize=True, return_levels=False, diag_cols=None,
                unit_weights=False, out=None):
        """kernels.py:430-433 / :455-469 / :471-476 in one launch.  `out`: preallocated (contiguous) result of the summed
        matrix (the all-gather input of parallel.sharded_K_symm)."""
        lib = _lib.load()
        dev = levels.device
        nl = levels.shape[0]
        n1 = levels.shape[1]
        n2 = levels.shape[2] if levels.dim() == 3 else 1
        w = torch.ones(nl, device=dev, dtype=torch.float32) if unit_weights else self._weights(dev)
        lev_out = torch.empty_like(levels) if return_levels else None
        if out is None or return_levels:
            out = None if return_levels else torch.empty(levels.shape[1:], device=dev, dtype=torch.float32)
        else:
            assert out.is_contiguous() and tuple(out.shape) == tuple(levels.shape[1:]) and out.dtype == torch.float32
        sym = bool(symmetric and normalize)
        with torch.cuda.device(dev):
            rc = lib.gpsig_normalize_weight_sum(levels.data_ptr(), nl, n1, n2, _ptr(diag1) if normalize else 0,
                                                _ptr(diag2) if normalize else 0, _ptr(diag_cols) if normalize else 0,
                                                float(self.jitter), int(sym), w.data_ptr(), _ptr(lev_out), _ptr(out),
                                                _stream())
        _lib.check(rc, "gpsig_normalize_weight_sum")
        return lev_out if return_levels else out

    def _scale_tens(self, Z, tensors=False):
        """kernels.py:366-398 (inducing tensors: `tensors`) / :357-361 (sequences): Z x gamma / lengthscales on the last axis."""
        lib = _lib.load()
        Z = self._to_dev(Z).contiguous()
        inv_ls = self._inv_ls(Z.device, tensors)
        if inv_ls is None:
            return Z
        out = torch.empty_like(Z)
        d = Z.shape[-1]
        with torch.cuda.device(Z.device):
            rc = lib.gpsig_scale_features(Z.data_ptr(), Z.numel() // d, d, inv_ls.data_ptr(), inv_ls.numel(), out.data_ptr(),
                                          _stream())
        _lib.check(rc, "gpsig_scale_features")
        return out

    def _base_gram(self, A, B=None):
        """static-kernel Gram of already-scaled points (rows, d) -> (rowsA, rowsB)  (kernels.py:225-230)."""
        lib = _lib.load()
        r1, d = A.shape
        r2 = r1 if B is None else B.shape[0]
        out = torch.empty((r1, r2), device=A.device, dtype=torch.float32)
        keep, pptr = self._params_ptr()
        with torch.cuda.device(A.device):
            rc = lib.gpsig_gram(_KIND[self._kind], A.data_ptr(), r1, _ptr(B), r2, d, pptr, out.data_ptr(), r2, _stream())
        _lib.check(rc, "gpsig_gram")
        return out

    def _K_tens(self, Zs, increments=False):
        """kernels.py:263-283 on SCALED tensors; returns (M+1, nz, nz)."""
        lib = _lib.load()
        T, nz, d = Zs.shape[0], Zs.shape[1], Zs.shape[-1]
        rows = nz * (2 if increments else 1)
        M = torch.empty((T, rows, rows), device=Zs.device, dtype=torch.float32)
        Zf = Zs.reshape(T, rows, d)
        for k in range(T):
            M[k] = self._base_gram(Zf[k])
        out = torch.empty((self.num_levels + 1, nz, nz), device=Zs.device, dtype=torch.float32)
        with torch.cuda.device(Zs.device):
            rc = lib.gpsig_tensor_kern_levels(M.data_ptr(), self.num_levels, nz, nz, int(bool(increments)), out.data_ptr(),
                                              _stream())
        _lib.check(rc, "gpsig_tensor_kern_levels")
        return out

    def _K_tens_vs_seq(self, Z, X, increments=False):
        """kernels.py:313-340 on RAW tensors / sequences (scaling is fused); returns (M+1, nz, n)."""
        lib = _lib.load()
        Z = self._to_dev(Z, like=X).contiguous()
        nz = Z.shape[1]
        n, L, d = X.shape
        out = torch.empty((self.num_levels + 1, nz, n), device=X.device, dtype=torch.float32)
        inv_ls = self._inv_ls(X.device)
        if self.lengthscales is None and self.num_lags > 0:
            # the reference leaves the tensors alone without lengthscales but still weights the lagged copies of the
            # sequences (kernels.py:360-361 vs :374): scale X here, nothing inside the kernel
            X, inv_ls = self._scale_tens(X), None
        keep, pptr = self._params_ptr()
        with torch.cuda.device(X.device):
            rc = lib.gpsig_tens_seq_kern_levels(_KIND[self._kind], pptr, Z.data_ptr(), nz, int(bool(increments)), X.data_ptr(),
                                                n, L, d, _ptr(inv_ls), self.num_levels, self.order, int(self.difference),
                                                out.data_ptr(), _stream())
        _lib.check(rc, "gpsig_tens_seq_kern_levels")
        return out

    # ---- low-rank mode (kernels.py:239-261, :285-311; low_rank_calculations.py) ----
    # Deviation from the literal reference: K() hands UNSCALED sequences to _K_seq_lr_feat (kernels.py:425,
    # :448-449) while every other method scales first; here lengthscales / lags always apply.
    def _lr_seeds(self):
        return self.lr_rng.integers(0, np.iinfo(np.int32).max, size=(max(self.num_levels - 1, 1), 2))

    def _lr_samples(self, *point_sets):
        """kernels.py:444-446 / :562-563: Nystrom landmarks drawn from the union of all points of the call."""
        pts = torch.cat([p.reshape(-1, p.shape[-1]) for p in point_sets], dim=0)
        idx, _ = _lr._draw_indices(pts.shape[0], min(self.num_components, pts.shape[0]), self.lr_rng)
        return pts[torch.as_tensor(idx, device=pts.device)]

    def _K_seq_lr_feat(self, Xsc, nys_samples=None, seeds=None):
        """kernels.py:239-261 on SCALED sequences (n, L, d): list of low-rank factors per level."""
        n, L, d = Xsc.shape
        F = _lr.Nystrom_map(Xsc.reshape(n * L, d), self._base_gram, nys_samples, self.num_components, rng=self.lr_rng)
        return _algs.signature_kern_first_order_lr_feature(F.reshape(n, L, -1), self.num_levels, self.rank_bound, self.sparsity,
                                                           seeds, difference=self.difference)

    def _K_tens_lr_feat(self, Zsc, increments=False, nys_samples=None, seeds=None):
        """kernels.py:285-311 on SCALED tensors."""
        T, nz, d = Zsc.shape[0], Zsc.shape[1], Zsc.shape[-1]
        F = _lr.Nystrom_map(Zsc.reshape(-1, d), self._base_gram, nys_samples, self.num_components, rng=self.lr_rng)
        if increments:
            F = F.reshape(T, nz, 2, -1)
            F = F[:, :, 1, :] - F[:, :, 0, :]
        else:
            F = F.reshape(T, nz, -1)
        return _algs.tensor_kern_lr_feature(F, self.num_levels, self.rank_bound, self.sparsity, seeds)

    @staticmethod
    def _lr_gram(PhiA, PhiB):
        return torch.stack([a @ b.transpose(0, 1) for a, b in zip(PhiA, PhiB)], dim=0).contiguous()

    @staticmethod
    def _lr_diag(Phi):
        return torch.stack([torch.sum(p * p, dim=-1) for p in Phi], dim=0).contiguous()

    # ---- public API (kernels.py:400-761) ----
    def K(self, X, X2=None, presliced=False, return_levels=False, presliced_X=False, presliced_X2=False):
        """kernels.py:400-476."""
        self._check_supported()
        if presliced:
            presliced_X = presliced_X2 = True
        if self._grad_mode(X, X2):
            self._check_grad_supported()
            Xt = self._seqs_t(X, presliced_X)
            if X2 is None:
                lv = _ad.K_seq_levels(self, Xt)
                return _ad.finish(self, lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
            X2t = self._seqs_t(X2, presliced_X2)
            lv = _ad.K_seq_levels(self, Xt, X2t)
            d1 = d2 = None
            if self.normalization:
                d1, d2 = _ad.K_seq_diag_levels(self, Xt), _ad.K_seq_diag_levels(self, X2t)
            return _ad.finish(self, lv, d1, d2, normalize=self.normalization, return_levels=return_levels)
        Xs = self._seqs(X, presliced_X)
        if X2 is None:
            if self.low_rank:                                                                    # kernels.py:424-426
                Phi = self._K_seq_lr_feat(self._scale_tens(Xs))
                lv = self._lr_gram(Phi, Phi)
            else:
                lv = self._K_seq(Xs)
            return self._finish(lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
        X2s = self._seqs(X2, presliced_X2)
        if self.low_rank:                                                                        # kernels.py:442-451, :456-458
            Xc, X2c = self._scale_tens(Xs), self._scale_tens(X2s)
            seeds, nys = self._lr_seeds(), self._lr_samples(Xc, X2c)
            Phi, Phi2 = self._K_seq_lr_feat(Xc, nys, seeds), self._K_seq_lr_feat(X2c, nys, seeds)
            lv = self._lr_gram(Phi, Phi2)
            d1, d2 = (self._lr_diag(Phi), self._lr_diag(Phi2)) if self.normalization else (None, None)
            return self._finish(lv, d1, d2, normalize=self.normalization, return_levels=return_levels)
        lv = self._K_seq(Xs, X2s)
        d1 = d2 = None
        if self.normalization:
            d1, d2 = self._K_seq_diag(Xs), self._K_seq_diag(X2s)
        return self._finish(lv, d1, d2, normalize=self.normalization, return_levels=return_levels)

    def Kdiag(self, X, presliced=False, return_levels=False):
        """kernels.py:478-510."""
        self._check_supported()
        n = X.shape[0]
        if self.normalization and self._grad_mode():
            w = self._weights_tensor(self._dev(X)).to(torch.float32)
            return w[:, None].expand(-1, n) if return_levels else w.sum().expand(n)
        if self.normalization:
            dev = self._dev(X)
            w = self._weights(dev)
            if return_levels:
                return w[:, None].expand(-1, n).contiguous()
            return torch.full((n,), float(self.sigma * np.sum(self.variances)), device=dev, dtype=torch.float32)
        if self._grad_mode(X):
            self._check_grad_supported()
            lv = _ad.K_seq_diag_levels(self, self._seqs_t(X, presliced))
            return _ad.finish(self, lv[:, :, None], normalize=False, return_levels=return_levels).squeeze(-1)
        Xs = self._seqs(X, presliced)
        if self.low_rank:                                                                        # kernels.py:499-501
            lv = self._lr_diag(self._K_seq_lr_feat(self._scale_tens(Xs)))
        else:
            lv = self._K_seq_diag(Xs)
        return self._finish(lv[:, :, None].contiguous(), normalize=False, return_levels=return_levels).squeeze(-1)

    def K_tens(self, Z, return_levels=False, increments=False):
        """kernels.py:512-536."""
        self._check_supported()
        if self._grad_mode(Z):
            self._check_grad_supported()
            lv = _ad.K_tens_levels(self, self._tens_t(Z), increments)
            return _ad.finish(self, lv, normalize=False, return_levels=return_levels)
        if self.low_rank:                                                                        # kernels.py:525-527
            Phi = self._K_tens_lr_feat(self._scale_tens(Z, tensors=True), increments)
            lv = self._lr_gram(Phi, Phi)
        else:
            lv = self._K_tens(self._scale_tens(Z, tensors=True), increments)
        return self._finish(lv, normalize=False, return_levels=return_levels)

    def K_tens_vs_seq(self, Z, X, return_levels=False, increments=False, presliced=False):
        """kernels.py:538-588."""
        self._check_supported()
        if self._grad_mode(Z, X):
            self._check_grad_supported()
            Xt = self._seqs_t(X, presliced)
            lv = _ad.K_tens_vs_seq_levels(self, self._tens_t(Z, Xt), Xt, increments)
            d2 = _ad.K_seq_diag_levels(self, Xt) if self.normalization else None
            return _ad.finish(self, lv, None, d2, normalize=self.normalization, return_levels=return_levels)
        Xs = self._seqs(X, presliced)
        if self.low_rank:                                                                        # kernels.py:560-568, :573-574
            Zc, Xc = self._scale_tens(Z, tensors=True), self._scale_tens(Xs)
            seeds, nys = self._lr_seeds(), self._lr_samples(Zc, Xc)
            PhiZ, PhiX = self._K_tens_lr_feat(Zc, increments, nys, seeds), self._K_seq_lr_feat(Xc, nys, seeds)
            lv = self._lr_gram(PhiZ, PhiX)
            d2 = self._lr_diag(PhiX) if self.normalization else None
            return self._finish(lv, None, d2, normalize=self.normalization, return_levels=return_levels)
        lv = self._K_tens_vs_seq(Z, Xs, increments)
        d2 = self._K_seq_diag(Xs) if self.normalization else None
        return self._finish(lv, None, d2, normalize=self.normalization, return_levels=return_levels)

    def K_tens_n_seq_covs(self, Z, X, full_X_cov=False, return_levels=False, increments=False, presliced=False):
        """kernels.py:590-671."""
        self._check_supported()
        if self._grad_mode(Z, X):
            self._check_grad_supported()
            Xt = self._seqs_t(X, presliced)
            Zt = self._tens_t(Z, Xt)
            Kzz = _ad.finish(self, _ad.K_tens_levels(self, Zt, increments), normalize=False, return_levels=return_levels)
            Kzx_lv = _ad.K_tens_vs_seq_levels(self, Zt, Xt, increments)
            if full_X_cov:
                Kxx_lv = _ad.K_seq_levels(self, Xt)
                dg = torch.diagonal(Kxx_lv, dim1=1, dim2=2) if self.normalization else None       # kernels.py:632-638
                Kxx = _ad.finish(self, Kxx_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
                Kzx = _ad.finish(self, Kzx_lv, None, dg, normalize=self.normalization, return_levels=return_levels)
            else:
                dg = _ad.K_seq_diag_levels(self, Xt)
                Kzx = _ad.finish(self, Kzx_lv, None, dg, normalize=self.normalization, return_levels=return_levels)
                if self.normalization:                                                               # kernels.py:655-661
                    w = self._weights_tensor(Xt.device).to(torch.float32)
                    Kxx = w[:, None].expand(-1, Xt.shape[0]) if return_levels else w.sum().expand(Xt.shape[0])
                else:
                    Kxx = _ad.finish(self, dg[:, :, None], normalize=False, return_levels=return_levels).squeeze(-1)
            return Kzz, Kzx, Kxx
        Xs = self._seqs(X, presliced)
        PhiX = None
        if self.low_rank:                                                                        # kernels.py:612-621
            Zc, Xc = self._scale_tens(Z, tensors=True), self._scale_tens(Xs)
            seeds, nys = self._lr_seeds(), self._lr_samples(Zc, Xc)
            PhiZ, PhiX = self._K_tens_lr_feat(Zc, increments, nys, seeds), self._K_seq_lr_feat(Xc, nys, seeds)
            Kzz = self._finish(self._lr_gram(PhiZ, PhiZ), normalize=False, return_levels=return_levels)
            Kzx_lv = self._lr_gram(PhiZ, PhiX)
        else:
            Kzz = self._finish(self._K_tens(self._scale_tens(Z, tensors=True), increments), normalize=False, return_levels=return_levels)
            Kzx_lv = self._K_tens_vs_seq(Z, Xs, increments)
        if full_X_cov:
            Kxx_lv = self._lr_gram(PhiX, PhiX) if self.low_rank else self._K_seq(Xs)
            dg = Kxx_lv.diagonal(dim1=1, dim2=2).contiguous() if self.normalization else None       # kernels.py:632-638
            Kxx = self._finish(Kxx_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
            Kzx = self._finish(Kzx_lv, None, dg, normalize=self.normalization, return_levels=return_levels)
        else:
            dg = self._lr_diag(PhiX) if self.low_rank else self._K_seq_diag(Xs)
            Kzx = self._finish(Kzx_lv, None, dg, normalize=self.normalization, return_levels=return_levels)
            if self.normalization:                                                                   # kernels.py:655-661
                w = self._weights(Xs.device)
                Kxx = w[:, None].expand(-1, Xs.shape[0]).contiguous()
                if not return_levels:
                    Kxx = torch.full((Xs.shape[0],), float(self.sigma * np.sum(self.variances)), device=Xs.device,
                                     dtype=torch.float32)
            else:
                Kxx = self._finish(dg[:, :, None].contiguous(), normalize=False, return_levels=return_levels).squeeze(-1)
        return Kzz, Kzx, Kxx

    def K_seq_n_seq_covs(self, X, X2, full_X2_cov=False, return_levels=False, presliced=False, literal=False):
        """
        kernels.py:673-761 (InducingSequences).  X (inducing sequences) is never sliced (:679-680).  The reference
        divides Kxx2 by sqrt(diag Kxx) twice in the normalised diagonal branch (:713 then :750);
        literal=True reproduces that, the default divides once.  full_X2_cov with normalisation raises NameError in
        the reference (Q2); the evident intent is implemented here.
        """
        self._check_supported()
        if self._grad_mode(X, X2):
            self._check_grad_supported()
            Xa, Xb = self._seqs_t(X, presliced=True), self._seqs_t(X2, presliced)
            Kxx_lv, Kxx2_lv = _ad.K_seq_levels(self, Xa), _ad.K_seq_levels(self, Xa, Xb)
            d1 = torch.diagonal(Kxx_lv, dim1=1, dim2=2) if self.normalization else None
            Kxx = _ad.finish(self, Kxx_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
            if full_X2_cov:
                K22_lv = _ad.K_seq_levels(self, Xb)
                d2 = torch.diagonal(K22_lv, dim1=1, dim2=2) if self.normalization else None
                K22 = _ad.finish(self, K22_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
            else:
                d2 = _ad.K_seq_diag_levels(self, Xb)
                if self.normalization:
                    w = self._weights_tensor(Xb.device).to(torch.float32)
                    K22 = w[:, None].expand(-1, Xb.shape[0]) if return_levels else w.sum().expand(Xb.shape[0])
                else:
                    K22 = _ad.finish(self, d2[:, :, None], normalize=False, return_levels=return_levels).squeeze(-1)
            if self.normalization and literal and not full_X2_cov:                                   # quirk Q4
                once = _ad.finish(self, Kxx2_lv, d1, None, normalize=True, return_levels=True, unit_weights=True)
                Kxx2 = _ad.finish(self, once, d1, d2, normalize=True, return_levels=return_levels)
            else:
                Kxx2 = _ad.finish(self, Kxx2_lv, d1, d2, normalize=self.normalization, return_levels=return_levels)
            return Kxx, Kxx2, K22
        Xa = self._seqs(X, presliced=True)
        Xb = self._seqs(X2, presliced)
        Phi2 = None
        if self.low_rank:                                                                        # kernels.py:693-702
            Xac, Xbc = self._scale_tens(Xa), self._scale_tens(Xb)
            seeds, nys = self._lr_seeds(), self._lr_samples(Xac, Xbc)
            Phi, Phi2 = self._K_seq_lr_feat(Xac, nys, seeds), self._K_seq_lr_feat(Xbc, nys, seeds)
            Kxx_lv, Kxx2_lv = self._lr_gram(Phi, Phi), self._lr_gram(Phi, Phi2)
        else:
            Kxx_lv = self._K_seq(Xa)
            Kxx2_lv = self._K_seq(Xa, Xb)
        d1 = Kxx_lv.diagonal(dim1=1, dim2=2).contiguous() if self.normalization else None
        Kxx = self._finish(Kxx_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
        if full_X2_cov:
            K22_lv = self._lr_gram(Phi2, Phi2) if self.low_rank else self._K_seq(Xb)
            d2 = K22_lv.diagonal(dim1=1, dim2=2).contiguous() if self.normalization else None
            K22 = self._finish(K22_lv, symmetric=True, normalize=self.normalization, return_levels=return_levels)
        else:
            d2 = self._lr_diag(Phi2) if self.low_rank else self._K_seq_diag(Xb)
            if self.normalization:
                w = self._weights(Xb.device)
                K22 = w[:, None].expand(-1, Xb.shape[0]).contiguous()
                if not return_levels:
                    K22 = torch.full((Xb.shape[0],), float(self.sigma * np.sum(self.variances)), device=Xb.device,
                                     dtype=torch.float32)
            else:
                K22 = self._finish(d2[:, :, None].contiguous(), normalize=False, return_levels=return_levels).squeeze(-1)
        if self.normalization and literal and not full_X2_cov:
            # Q4: Kxx2 is divided by sqrt(diag Kxx + jitter) at :713 and again at :750 -- two normalisation launches
            once = self._finish(Kxx2_lv, d1, None, normalize=True, return_levels=True, unit_weights=True)
            Kxx2 = self._finish(once, d1, d2, normalize=True, return_levels=return_levels)
        else:
            Kxx2 = self._finish(Kxx2_lv, d1, d2, normalize=self.normalization, return_levels=return_levels)
        return Kxx, Kxx2, K22

    # ---- numpy-facing helpers (kernels.py:141-186) ----
    @staticmethod
    def _np(t):
        return t.detach().cpu().numpy()

    def compute_K(self, X, Y):
        return self._np(self.K(X, Y))

    def compute_K_symm(self, X):
        return self._np(self.K(X))

    def compute_base_kern_symm(self, X):
        """kernels.py:150-156: (N, N, L, L) static-kernel Gram of the scaled sequences."""
        self._check_supported()
        Xs = self._seqs(X, presliced=True)
        n, L, d = Xs.shape
        flat = self._scale_tens(Xs.reshape(1, n * L, d)).reshape(n * L, d)
        M = self._base_gram(flat).reshape(n, L, n, L).permute(0, 2, 1, 3)
        return self._np(M)

    def compute_K_level_diags(self, X):
        return self._np(self.Kdiag(X, return_levels=True))

    def compute_K_levels(self, X, X2):
        return self._np(self.K(X, X2, return_levels=True))

    def compute_Kdiag(self, X):
        return self._np(self.Kdiag(X))

    def compute_K_tens(self, Z):
        return self._np(self.K_tens(Z, return_levels=False))

    def compute_K_tens_vs_seq(self, Z, X):
        return self._np(self.K_tens_vs_seq(Z, X, return_levels=False))

    def compute_K_incr_tens(self, Z):
        return self._np(self.K_tens(Z, increments=True, return_levels=False))

    def compute_K_incr_tens_vs_seq(self, Z, X):
        return self._np(self.K_tens_vs_seq(Z, X, increments=True, return_levels=False))


class SignatureLinear(SignatureKernel):
    """kernels.py:786-806."""
    _kind = "linear"


class SignatureCosine(SignatureKernel):
    """kernels.py:808-828."""
    _kind = "cosine"


class SignaturePoly(SignatureKernel):
    """kernels.py:831-848."""
    _kind = "poly"

    def __init__(self, input_dim, num_features, num_levels, gamma=1, degree=3, **kwargs):
        super().__init__(input_dim, num_features, num_levels, **kwargs)
        # the reference stores the offset as `self.gamma` (kernels.py:838), which overwrites the lag weights of the base
        # class (:82) and breaks num_lags > 0; here it has its own name, `gamma` stays the lag weights
        self.gamma_poly = float(gamma)
        self.degree = float(degree)
        if self.num_lags == 0:
            self.gamma = self.gamma_poly   # same attribute name as the reference when there is no clash

    def _static_params(self):
        return [self.gamma_poly, self.degree]


class SignatureRBF(SignatureKernel):
    """kernels.py:850-864."""
    _kind = "rbf"


SignatureGauss = SignatureRBF  # kernels.py:868


class SignatureMix(SignatureKernel):
    """kernels.py:870-892."""
    _kind = "mix"

    def __init__(self, input_dim, num_features, num_levels, **kwargs):
        super().__init__(input_dim, num_features, num_levels, **kwargs)
        self.mixing = 0.5                                                                          # kernels.py:876

    def _static_params(self):
        return [self.mixing]


class SignatureSpectral(SignatureKernel):
    """kernels.py:894-942.  Families 'gauss' / 'exp'; 'mixed' references an undefined name upstream."""
    _kind = "spectral"

    def __init__(self, input_dim, num_features, num_levels, family='gauss', Q=5, **kwargs):
        kwargs.pop("lengthscales", None)
        super().__init__(input_dim, num_features, num_levels, lengthscales=None, **kwargs)
        if family in ('exp', 'exponential'):
            self.family = 'exp'
        elif family in ('gauss', 'gaussian', 'rbf'):
            self.family = 'rbf'
        elif family in ('mixed', 'mix'):
            raise NotImplementedError("the 'mixed' spectral family is broken in the reference (kernels.py:932 uses an undefined Q)")
        else:
            raise ValueError("Unrecognized spectral family name.")
        if Q > 8 or self.num_features * (self.num_lags + 1) > 16:
            raise NotImplementedError("the device spectral kernel supports Q <= 8 and at most 16 (lagged) features")
        self.Q = int(Q)
        lag_gamma = getattr(self, "gamma", None)          # the lag weights of the base class share the reference's attribute name
        self.alpha = np.exp(np.random.randn(Q))                                                   # kernels.py:913-915
        self.omega = np.exp(np.random.randn(Q, self.num_features))
        self.spec_gamma = np.exp(np.random.randn(Q, self.num_features))
        if lag_gamma is not None:
            self.gamma = lag_gamma

    def _static_params(self):
        reps = self.num_lags + 1
        om = np.tile(np.asarray(self.omega, dtype=np.float64), (1, reps))
        ga = np.tile(np.asarray(self.spec_gamma, dtype=np.float64), (1, reps))
        return [1.0 if self.family == 'exp' else 0.0, float(self.Q), float(om.shape[1])] + list(np.asarray(self.alpha).ravel()) \
            + list(om.ravel()) + list(ga.ravel())


class SignatureMatern12(SignatureKernel):
    """kernels.py:944-958."""
    _kind = "matern12"


class SignatureMatern32(SignatureKernel):
    """kernels.py:964-977."""
    _kind = "matern32"


class SignatureMatern52(SignatureKernel):
    """kernels.py:981-993."""
    _kind = "matern52"


SignatureLaplace = SignatureMatern12      # kernels.py:961
SignatureExponential = SignatureMatern12  # kernels.py:962
