import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from gpsig_b200 import kernels, _lib
from oracle import gpsig_oracle as O
from util import random_walks
for (n, L, d, M) in ((21, 45, 3, 4), (21, 45, 3, 3), (8, 45, 3, 4), (21, 40, 3, 4), (21, 45, 8, 4), (21, 64, 3, 4)):
    X = random_walks(n, L, d, L + d).reshape(n, -1)
    ls = 0.5 * np.sqrt(d) + 0.5
    k = kernels.SignatureRBF(L * d, d, M, lengthscales=ls, normalization=False)
    ko = O.SignatureKernelOracle("rbf", L * d, d, M, lengthscales=ls, normalization=False)
    ref = ko.K(X, return_levels=True)
    got = k.K(X, return_levels=True).cpu().numpy()
    _lib.set_knob("warpfused", 0)
    pipe = k.K(X, return_levels=True).cpu().numpy()
    _lib.set_knob("warpfused", 1)
    for m in range(1, M + 1):
        e = np.abs(got[m] - ref[m]) / np.max(np.abs(ref[m]))
        ep = np.abs(pipe[m] - ref[m]) / np.max(np.abs(ref[m]))
        bad = np.argwhere(e > 1e-4)
        print((n, L, d, M), "level", m, "fused err %.2e pipe err %.2e" % (e.max(), ep.max()), "bad entries", len(bad), bad[:12].tolist())
