"""tcgen05 Kuf kernel (tens_tc.cu) against the CUDA-core kernel and the oracle on small problems."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpsig_b200 import kernels, _lib
from oracle import gpsig_oracle as O
from util import random_walks
import bench
for (nz, n, L, d, M) in ((10, 6, 128, 8, 5), (3, 2, 64, 3, 2), (20, 37, 100, 10, 6), (9, 5, 45, 4, 3), (64, 200, 128, 8, 5)):
    X = random_walks(n, L, d, 1).reshape(n, -1)
    Z = bench.synth_Z(X, L, d, M, nz)
    ls = float(np.sqrt(d))
    k = kernels.SignatureRBF(L * d, d, M, lengthscales=ls, normalization=False)
    ko = O.SignatureKernelOracle("rbf", L * d, d, M, lengthscales=ls, normalization=False)
    _lib.set_knob("tens_tc", 0)
    ref_fast = k.K_tens_vs_seq(Z, X, increments=True, return_levels=True).cpu().numpy()
    _lib.set_knob("tens_tc", 1)
    t0 = time.time()
    got = k.K_tens_vs_seq(Z, X, increments=True, return_levels=True)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    dt = time.time() - t0
    ref = ko.K_tens_vs_seq(Z, X, increments=True, return_levels=True) if nz * n <= 1000 else ref_fast.astype(np.float64)
    errs = ["%.1e/%.1e" % (np.abs(got[m] - ref[m]).max() / max(np.abs(ref[m]).max(), 1e-30), np.abs(ref_fast[m] - ref[m]).max() / max(np.abs(ref[m]).max(), 1e-30)) for m in range(M + 1)]
    print((nz, n, L, d, M), "%.3fs" % dt, "err tc/fast per level:", errs, flush=True)
