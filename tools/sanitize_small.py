"""Small invocations of every tuned kernel for compute-sanitizer (memcheck / racecheck): warp-fused, two-kernel pipeline,
operator-level TMA kernel, Kuf kernel.  Usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsig_b200 import kernels, _lib, signature_algs as S  # noqa: E402

rng = np.random.default_rng(0)
n, L, d, M = 10, 64, 4, 3
X = (np.cumsum(rng.standard_normal((n, L, d)), axis=1) / np.sqrt(L)).reshape(n, -1)
Y = (np.cumsum(rng.standard_normal((5, L, d)), axis=1) / np.sqrt(L)).reshape(5, -1)
Z = 0.4 * rng.standard_normal((M * (M + 1) // 2, 6, 2, d))
for cls in (kernels.SignatureLinear, kernels.SignatureRBF):
    k = cls(L * d, d, M, lengthscales=1.3)
    for knob in (1, 0):
        _lib.set_knob("warpfused", knob)
        a, b = k.K(X), k.K(X, Y)
        torch.cuda.synchronize()
    c = k.K_tens_vs_seq(Z, X, increments=True)
    torch.cuda.synchronize()
# round 2: ragged strips (L = 45: partial and past-the-end strips of the anchored form), d = 8 / M = 5 (the cfg4
# instantiation), diagonal mode; the tcgen05 Kuf kernel in its three operand layouts (d = 8: one atom; d = 10: ZOUT, with a
# partial last tile at L = 100; d = 12: two atoms); the reverse-mode kernels; the higher-order warp kernel
n2, L2, d2, M2 = 7, 45, 8, 5
X2 = (np.cumsum(rng.standard_normal((n2, L2, d2)), axis=1) / np.sqrt(L2)).reshape(n2, -1)
k2 = kernels.SignatureRBF(L2 * d2, d2, M2, lengthscales=2.0)
e, f = k2.K(X2), k2.Kdiag(X2)
torch.cuda.synchronize()
for dd, LL, MM in ((8, 128, 5), (10, 100, 6), (12, 70, 4)):
    Xk = (np.cumsum(rng.standard_normal((9, LL, dd)), axis=1) / np.sqrt(LL)).reshape(9, -1)
    Zk = Xk.reshape(9, LL, dd)[rng.integers(0, 9, size=(MM * (MM + 1) // 2, 11)), rng.integers(0, LL - 1, size=(MM * (MM + 1) // 2, 11))]
    Zk = np.stack([Zk, Zk + 0.3 * rng.standard_normal(Zk.shape)], axis=2)
    kk = kernels.SignatureRBF(LL * dd, dd, MM, lengthscales=float(np.sqrt(dd)))
    g = kk.K_tens_vs_seq(Zk, Xk, increments=True)
    torch.cuda.synchronize()
k3 = kernels.SignatureLinear(16 * 3, 3, 3, order=2)
h = k3.K((np.cumsum(rng.standard_normal((6, 16, 3)), axis=1) / 4.0).reshape(6, -1))
torch.cuda.synchronize()
from gpsig_b200 import autodiff as AD  # noqa: E402
Dt = (0.1 * torch.randn((4, 20, 5, 24), device="cuda", dtype=torch.float64)).requires_grad_(True)
AD.sigkern_first_order(Dt, 3).sum().backward()
Ht = (0.1 * torch.randn((6, 4, 5, 18), device="cuda", dtype=torch.float64)).requires_grad_(True)
try:
    AD.tens_vs_seq_first_order(Ht, 3).sum().backward()
except Exception as ex:  # signature differences are not what this script is about
    print("tens_vs_seq_first_order skipped:", ex)
torch.cuda.synchronize()
G = torch.randn((3, 64, 5, 64), device="cuda")
S.signature_kern_first_order(G, 3, difference=True)
S.signature_kern_first_order(G, 3, difference=False)
torch.cuda.synchronize()
print("sanitize_small ok", float(a.sum()), float(b.sum()), float(c.sum()), float(e.sum()), float(f.sum()), float(g.sum()), float(h.sum()))
