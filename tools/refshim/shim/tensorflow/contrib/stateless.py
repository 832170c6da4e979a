"""
tensorflow.contrib.stateless stand-in: seed-keyed deterministic draws from numpy (NOT bit-compatible with TF's
Philox stream -- TF is absent, so low-rank parity is defined on injected draws; every draw is logged).
"""
import numpy as np
from .. import Tensor, _a, _shape_arg, _log


def _gen(seed, salt):
    seed = np.asarray(_a(seed)).astype(np.int64).ravel().tolist()
    return np.random.default_rng([salt] + [int(s) & 0x7FFFFFFF for s in seed])


def stateless_random_uniform(shape, seed, dtype=np.float64):
    return Tensor(_log("sl_uniform", _gen(seed, 1).uniform(size=_shape_arg(shape)).astype(dtype)))


def stateless_random_normal(shape, seed, dtype=np.float64):
    return Tensor(_log("sl_normal", _gen(seed, 2).standard_normal(_shape_arg(shape)).astype(dtype)))
