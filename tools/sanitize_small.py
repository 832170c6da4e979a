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
G = torch.randn((3, 64, 5, 64), device="cuda")
S.signature_kern_first_order(G, 3, difference=True)
S.signature_kern_first_order(G, 3, difference=False)
torch.cuda.synchronize()
print("sanitize_small ok", float(a.sum()), float(b.sum()), float(c.sum()))
