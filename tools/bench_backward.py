#!/usr/bin/env python
"""Timing of one SVGP training step (ELBO forward through the differentiable route + backward through the CUDA VJP kernels
+ Adam update) at minibatch sizes, next to the forward-only fast path on the same problem.  Diagnostic, not the bench."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gpsig_b200 import kernels, models, inducing_variables as iv, _lib  # noqa: E402

dev = torch.device("cuda:0")
for (N, L, d, M, nz) in ((256, 64, 8, 4, 64), (512, 100, 10, 5, 128), (1024, 128, 8, 5, 256)):
    Xnp = bench.synth_X(N, L, d).astype(np.float32)
    Znp = bench.synth_Z(Xnp.astype(np.float64), L, d, M, nz).astype(np.float32)
    X, Z = torch.from_numpy(Xnp).to(dev), torch.from_numpy(Znp).to(dev)
    Y = torch.from_numpy((np.arange(N)[:, None] % 2).astype(np.float64)).to(dev)
    kern = kernels.SignatureRBF(L * d, d, M, lengthscales=bench.lengthscales_for("rbf", d) * np.ones(d))
    model = models.SVGP(X, Y, kern, models.Bernoulli(), iv.InducingTensors(Z, M, increments=True), num_latent=1,
                        q_mu=0.1 * np.random.default_rng(7).standard_normal((nz, 1)))
    with torch.no_grad():
        model.compute_log_likelihood(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fwd = model.compute_log_likelihood()
        torch.cuda.synchronize()
        ms_fwd = (time.perf_counter() - t0) / 5 * 1e3
    model.set_trainable()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.training_loss()
        loss.backward()
        opt.step()
        return loss

    step(); torch.cuda.synchronize()
    lib = _lib.load()
    lib.gpsig_profile_reset(); lib.gpsig_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(5):
        loss = step()
    torch.cuda.synchronize()
    ms_step = (time.perf_counter() - t0) / 5 * 1e3
    lib.gpsig_profile_enable(0)
    prof = bench.read_profile(lib, _lib, (("vjp", 7), ("recursion", 2), ("recursion_other", 3), ("tens", 5), ("fused", 6)))
    print(json.dumps({"config": "SVGP training step N=%d L=%d d=%d M=%d Z=%d (RBF, Bernoulli)" % (N, L, d, M, nz),
                      "forward_only_fast_path_ms": ms_fwd, "training_step_ms": ms_step, "elbo_fast": fwd,
                      "elbo_differentiable": -float(loss.item()),
                      "kernel_ms_per_step": {k: v[0] / 5 for k, v in prof.items() if v[0] > 0},
                      "peak_memory_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del model, opt
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
