"""Shared helpers for the GPU parity tests."""
import os

import numpy as np

TOL = 1e-4  # north star: max|K_ours - K_ref| / max|K_ref| < 1e-4 (fp32 device arithmetic vs the fp64 reference)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    den = np.max(np.abs(ref))
    return float(np.max(np.abs(got - ref)) / (den if den > 0 else 1.0))


def assert_close(got, ref, tol=TOL, msg="", atol=0.0):
    """max|got - ref| < tol * max|ref| + atol.  `atol` (default 0) is for quantities that are sums of O(1) terms cancelling
    to (numerically) nothing in the fp64 reference -- fp32 cannot follow a level that is 1e-20 of its own increments."""
    got64, ref64 = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got64.shape == ref64.shape, (got64.shape, ref64.shape)
    err, den = float(np.max(np.abs(got64 - ref64))), float(np.max(np.abs(ref64)))
    assert np.isfinite(err) and err <= tol * den + atol, \
        "%s: max|err| = %.3e, max|ref| = %.3e: %.3e >= %.1e (atol %.1e)" % (msg, err, den, err / (den if den > 0 else 1.0), tol, atol)


def assert_levels_close(got, ref, tol=TOL, msg="", atol=0.0):
    """Level stacks: every level is held to the tolerance against its own scale."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    for m in range(ref.shape[0]):
        assert_close(got[m], ref[m], tol, "%s level %d" % (msg, m), atol)


def random_walks(n, L, d, seed):
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.standard_normal((n, L, d)), axis=1) / np.sqrt(L)
