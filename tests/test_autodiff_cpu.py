"""
Host-side plumbing of the differentiable route that needs no GPU: the static-kernel Gram blocks are evaluated chunk by chunk
under activation checkpointing (autodiff._maybe_checkpoint) -- values and gradients must equal the plain composition.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpsig_b200 import autodiff as AD  # noqa: E402


class _Kern:
    def __init__(self, kind):
        self._kind, self.difference, self.num_levels = kind, True, 3


def _increments(kern, Zf, X, T, nz, chunk):
    n, L, d = X.shape

    def block(Zf, Xc):
        nc = Xc.shape[0]
        M = AD.base_gram(kern, Zf, Xc.reshape(nc * L, d)).reshape(T, nz, 2, nc, L)
        M = M[:, :, 1] - M[:, :, 0]
        return (M[..., 1:] - M[..., :-1]).to(torch.float32)

    if chunk is None:
        return block(Zf, X)
    return torch.cat([AD._maybe_checkpoint(block, Zf, X[c:c + chunk]) for c in range(0, n, chunk)], dim=2)


def test_checkpointed_gram_blocks_match_the_plain_composition():
    torch.manual_seed(0)
    T, nz, d, n, L = 6, 3, 2, 5, 7
    for kind in ("rbf", "linear", "matern32"):
        kern = _Kern(kind)
        Zf = torch.randn(T * nz * 2, d, dtype=torch.float64, requires_grad=True)
        X = torch.randn(n, L, d, dtype=torch.float64, requires_grad=True)
        w = torch.randn(T, nz, n, L - 1)
        a = (_increments(kern, Zf, X, T, nz, None) * w).sum()
        ga = torch.autograd.grad(a, (Zf, X))
        b = (_increments(kern, Zf, X, T, nz, 2) * w).sum()
        gb = torch.autograd.grad(b, (Zf, X))
        assert abs(a.item() - b.item()) <= 1e-6 * max(1.0, abs(a.item())), kind
        for x, y in zip(ga, gb):
            assert torch.allclose(x, y, rtol=1e-12, atol=1e-12), kind


def test_no_checkpoint_without_gradients():
    calls = []

    def fn(t):
        calls.append(torch.is_grad_enabled())
        return t * 2

    t = torch.ones(3)
    assert torch.equal(AD._maybe_checkpoint(fn, t), t * 2) and len(calls) == 1
    with torch.no_grad():
        AD._maybe_checkpoint(fn, t.requires_grad_(True))
    assert len(calls) == 2
