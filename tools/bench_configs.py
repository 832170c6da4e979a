#!/usr/bin/env python
"""
Timings of BASELINE.json's other configurations (bench.py carries the headline one):

  cfg3  Kuf with InducingTensors Z=256, N=4096 L=128 d=8 M=5          (kernels.K_tens_vs_seq, increments on / off)
  cfg5  SVGP ELBO step, SignatureRBF N=8192 Z=512 L=100 d=10 M=6      (models.SVGP.compute_log_likelihood)

    python tools/bench_configs.py [--cfg cfg3|cfg5|all] [--steps K] [--scale s]

Each line is one JSON object: CUDA-event ms per call (median of K after 3 warm-ups), pairs/s, the per-class device times
from the library's profiling hooks, and a parity spot check against the fp64 oracle on a sub-block.  Not the bench
contract; results are copied to profiles/.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gpsig_b200 import kernels, inducing_variables as iv, models, _lib  # noqa: E402
from oracle import gpsig_oracle as O  # noqa: E402

CLASSES = (("prep", 0), ("producer", 1), ("recursion", 2), ("recursion_other", 3), ("epilogue", 4), ("tens", 5))


def walks(n, L, d, seed):
    rng = np.random.default_rng(seed)
    return (np.cumsum(rng.standard_normal((n, L, d)), axis=1) / np.sqrt(L)).reshape(n, L * d)


def timed(fn, steps):
    lib = _lib.load()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.gpsig_profile_reset()
    lib.gpsig_profile_enable(1)
    ms = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    lib.gpsig_profile_enable(0)
    prof = {}
    for name, c in CLASSES:
        t, n, u = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.gpsig_profile_read(c, ctypes.byref(t), ctypes.byref(n), ctypes.byref(u))
        if n.value:
            prof[name] = {"ms_per_call": t.value / steps, "launches_per_call": n.value / steps}
    lib.gpsig_profile_reset()
    return statistics.median(ms), prof


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def cfg3(args):
    N, L, d, M, Z = int(4096 * args.scale), 128, 8, 5, int(256 * args.scale)
    T = M * (M + 1) // 2
    rng = np.random.default_rng(2)
    X = torch.as_tensor(walks(N, L, d, 0), dtype=torch.float32).cuda()
    for kind, cls in (("rbf", kernels.SignatureRBF), ("linear", kernels.SignatureLinear)):
        ls = float(np.sqrt(d)) if kind == "rbf" else 1.0
        k = cls(L * d, d, M, lengthscales=ls)
        ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=ls)
        for inc in (True, False):
            Zt = 0.4 * rng.standard_normal((T, Z, 2, d) if inc else (T, Z, d))
            ms, prof = timed(lambda: k.K_tens_vs_seq(Zt, X, increments=inc), args.steps)
            got = k.K_tens_vs_seq(Zt[:, :16], X[:24], increments=inc).cpu().numpy()
            ref = ko.K_tens_vs_seq(Zt[:, :16], X[:24].cpu().numpy().astype(np.float64), increments=inc)
            bytes_alg = 4.0 * T * L * (2 if inc else 1) + 4 * (M + 1)
            print(json.dumps({"config": "cfg3 Kuf Z=%d N=%d L=%d d=%d M=%d %s increments=%s" % (Z, N, L, d, M, kind, inc),
                              "ms": ms, "zn_pairs_per_s": Z * N / (ms * 1e-3),
                              "equiv_GBps_if_gram_were_read": Z * N * bytes_alg / (ms * 1e-3) / 1e9,
                              "stages": prof, "relerr_vs_oracle_subblock": relerr(got, ref)}), flush=True)


def cfg5(args):
    N, L, d, M, Z = int(8192 * args.scale), 100, 10, 6, int(512 * args.scale)
    T = M * (M + 1) // 2
    rng = np.random.default_rng(3)
    Xn = walks(N, L, d, 0)
    X = torch.as_tensor(Xn, dtype=torch.float32).cuda()
    Yn = (rng.standard_normal((N, 1)) > 0).astype(np.float64)
    Y = torch.as_tensor(Yn, dtype=torch.float32).cuda()
    Zt = 0.4 * rng.standard_normal((T, Z, 2, d))
    q_mu = 0.3 * rng.standard_normal((Z, 1))
    q_sqrt = np.tril(0.1 * rng.standard_normal((1, Z, Z))) + np.eye(Z)[None]
    ls = float(np.sqrt(d))
    k = kernels.SignatureRBF(L * d, d, M, lengthscales=ls)
    feat = iv.InducingTensors(Zt, M, increments=True)
    m = models.SVGP(X, Y, k, models.Bernoulli(), feat, num_latent=1, q_mu=q_mu, q_sqrt=q_sqrt)
    ms, prof = timed(lambda: m._build_likelihood(X, Y), args.steps)
    elbo = m.compute_log_likelihood(X, Y)
    # parity on a sub-problem the oracle finishes in seconds
    ns, zs = 48, 12
    ko = O.SignatureKernelOracle("rbf", L * d, d, M, lengthscales=ls)
    qs = np.tril(q_sqrt[:, :zs, :zs])
    ms_ = models.SVGP(X[:ns], Y[:ns], k, models.Bernoulli(), iv.InducingTensors(Zt[:, :zs], M, increments=True),
                      num_latent=1, q_mu=q_mu[:zs], q_sqrt=qs)
    got = ms_.compute_log_likelihood()
    ref = O.svgp_elbo(ko, Zt[:, :zs], Xn[:ns], Yn[:ns], q_mu[:zs], qs, likelihood="bernoulli", increments=True)[0]
    print(json.dumps({"config": "cfg5 SVGP ELBO step SignatureRBF N=%d Z=%d L=%d d=%d M=%d R=1 Bernoulli" % (N, Z, L, d, M),
                      "ms": ms, "sequences_per_s": N / (ms * 1e-3), "elbo": elbo, "stages": prof,
                      "elbo_relerr_vs_oracle_subproblem": abs(got - ref) / abs(ref)}), flush=True)


def op(args):
    """operator level: gpsig_sigkern_levels on a caller-supplied Gram tensor [n1, L, n2, L] (the tensor-map TMA kernel)."""
    from gpsig_b200 import signature_algs as S
    n1, n2, L, M = int(30 * args.scale) or 1, 4096, 128, 5
    G = torch.randn((n1, L, n2, L), device="cuda", dtype=torch.float32) * 0.05
    for diff in (True, False):
        ms, prof = timed(lambda: S.signature_kern_first_order(G, M, difference=diff), args.steps)
        r = prof.get("recursion", {"ms_per_call": ms})
        gb = n1 * n2 * (4.0 * L * L + 4 * (M + 1)) / 1e9
        # parity of a corner against the oracle
        got = S.signature_kern_first_order(G[:2, :, :6, :].contiguous(), M, difference=diff).cpu().numpy()
        ref = O.signature_kern_first_order(G[:2, :, :6, :].cpu().numpy().astype(np.float64), M, difference=diff)
        print(json.dumps({"config": "operator signature_kern_first_order M[%d,%d,%d,%d] levels=%d difference=%s" % (n1, L, n2, L, M, diff),
                          "ms": ms, "kernel_ms": r["ms_per_call"], "GBps": gb / (r["ms_per_call"] * 1e-3),
                          "frac_of_6550": gb / (r["ms_per_call"] * 1e-3) / 6550.4,
                          "relerr_vs_oracle_corner": max(relerr(got[m], ref[m]) for m in range(1, M + 1))}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="all")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="scale N and Z (quick runs)")
    a = ap.parse_args()
    if a.cfg in ("cfg3", "all"):
        cfg3(a)
    if a.cfg in ("cfg5", "all"):
        cfg5(a)
    if a.cfg in ("op", "all"):
        op(a)
