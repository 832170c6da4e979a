#!/usr/bin/env python
"""Where the cfg5 SVGP ELBO step spends its time outside the Kuf kernel: wall-clock per phase with a device synchronise on
both sides (diagnostic, not a bench)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gpsig_b200 import kernels, models, inducing_variables as iv, settings  # noqa: E402
from gpsig_b200.models import base_conditional, gauss_kl  # noqa: E402

wl = bench.WORKLOADS["cfg5"]
N, L, d, M, nz = wl["N"], wl["L"], wl["d"], wl["M"], wl["Z"]
dev = torch.device("cuda:0")
X = torch.from_numpy(bench.synth_X(N, L, d).astype(np.float32)).to(dev)
Z = torch.from_numpy(bench.synth_Z(X.cpu().numpy().astype(np.float64), L, d, M, nz).astype(np.float32)).to(dev)
kern = kernels.SignatureRBF(L * d, d, M, lengthscales=bench.lengthscales_for("rbf", d))
Y = torch.from_numpy((np.arange(N)[:, None] % 2).astype(np.float64)).to(dev)
feat = iv.InducingTensors(Z, M, increments=True)
model = models.SVGP(X, Y, kern, models.Bernoulli(), feat, num_latent=1, q_mu=0.1 * np.random.default_rng(7).standard_normal((nz, 1)))


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


ms_all, _ = t(model.compute_log_likelihood)
ms_cov, (Kzz, Kzx, Kxx) = t(lambda: iv.Kuu_Kuf_Kff(feat, kern, X, jitter=settings.jitter))
ms_kuf, _ = t(lambda: iv.Kuf(feat, kern, X))
ms_kuu, _ = t(lambda: iv.Kuu(feat, kern, jitter=settings.jitter))
ms_diag, _ = t(lambda: kern.Kdiag(X))
Kzz64, Kzx64, Kxx64 = Kzz.double(), Kzx.double(), Kxx.double()
ms_cast, _ = t(lambda: (Kzz.double(), Kzx.double(), Kxx.double()))
ms_up, (qm, qs) = t(lambda: (model._dev(model.q_mu, dev), torch.tril(model._dev(model.q_sqrt, dev))))
ms_cond, (fm, fv) = t(lambda: base_conditional(Kzx64, Kzz64, Kxx64, qm, full_cov=False, q_sqrt=qs, white=True))
ms_kl, _ = t(lambda: gauss_kl(qm, qs, K=None))
ms_ve, _ = t(lambda: model.likelihood.variational_expectations(fm, fv, Y).sum())
print("ELBO step %.3f ms | Kuu_Kuf_Kff %.3f (Kuf %.3f, Kuu %.3f, Kdiag %.3f) | fp64 casts %.3f | q upload (x2 per step) %.3f | "
      "base_conditional %.3f | gauss_kl %.3f | variational expectations %.3f" %
      (ms_all, ms_cov, ms_kuf, ms_kuu, ms_diag, ms_cast, ms_up, ms_cond, ms_kl, ms_ve))
