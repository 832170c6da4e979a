import numpy as np
float_type = np.float64
int_type = np.int32
jitter = 1e-6  # GPflow 1.5.1 default settings.numerics.jitter_level


class _N:
    jitter_level = 1e-6


class _D:
    float_type = np.float64
    int_type = np.int32


numerics = _N()
dtypes = _D()
