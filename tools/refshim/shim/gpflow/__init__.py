"""
Minimal GPflow 1.5.1 stand-in (TEST INFRASTRUCTURE ONLY) -- just enough surface for the reference's
gpsig/kernels.py, inducing_variables.py, low_rank_calculations.py and lags.py to import and run eagerly on the
numpy-backed tensorflow shim.  Parameters are plain constrained-space values (no transforms are applied because the
reference only ever reads the constrained value inside @params_as_tensors).
"""
from . import settings, transforms, params, decors, kernels, features, dispatch, conditionals, kullback_leiblers  # noqa: F401
