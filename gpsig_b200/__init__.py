"""
gpsig_b200 -- B200-native (sm_100a) signature-kernel covariance path behind the tgcsaba/GPSig API.

Module layout mirrors gpsig/__init__.py:1-6 of the reference: `kernels`, `inducing_variables`, `signature_algs`,
`models` (plus `parallel` for the multi-GPU row-sharded covariance, which the reference does not have).
All arithmetic on the path runs in the CUDA library built from gpsig_b200/csrc (C ABI: include/gpsig_b200.h);
torch tensors are device containers.
"""
from . import settings  # noqa: F401
from . import low_rank_calculations  # noqa: F401
from . import signature_algs  # noqa: F401
from . import kernels  # noqa: F401
from . import inducing_variables  # noqa: F401
from . import models  # noqa: F401
from . import parallel  # noqa: F401

__version__ = "0.1.0"
