import numpy as np
import tensorflow as tf


def Parameter(value, transform=None, dtype=None, trainable=True, **kw):
    """constrained-space value as an eager shim tensor"""
    return tf.Tensor(np.array(tf._a(value), dtype=dtype if dtype is not None else np.float64))


class ParamList(list):
    pass


def DataHolder(x):
    return tf.Tensor(x)


Minibatch = DataHolder
