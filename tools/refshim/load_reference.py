"""
Import the UNMODIFIED reference modules from /root/reference/gpsig on top of the numpy-backed tensorflow/gpflow shims.
Test infrastructure only (golden-vector generation in the build container; /root/reference does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

REF = os.environ.get("GPSIG_REFERENCE", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def load():
    if not os.path.isdir(os.path.join(REF, "gpsig")):
        raise RuntimeError("reference tree not found at %s" % REF)
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    if "gpsig" not in sys.modules:
        pkg = types.ModuleType("gpsig")          # bypass gpsig/__init__.py (it pulls models/training -> full gpflow)
        pkg.__path__ = [os.path.join(REF, "gpsig")]
        sys.modules["gpsig"] = pkg
    mods = {}
    for name in ("low_rank_calculations", "signature_algs", "lags", "kernels", "inducing_variables"):
        mods[name] = importlib.import_module("gpsig." + name)
    import tensorflow as tf
    return mods, tf
