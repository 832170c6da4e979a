import tensorflow as tf


class Kernel:
    """GPflow 1.5.1 Kernel base: input_dim / active_dims bookkeeping and _slice (gpflow/kernels.py in 1.5.1)."""

    def __init__(self, input_dim, active_dims=None, name=None):
        self.input_dim = int(input_dim)
        if active_dims is None:
            self.active_dims = slice(self.input_dim)
        elif isinstance(active_dims, slice):
            self.active_dims = active_dims
        else:
            self.active_dims = [int(a) for a in active_dims]
        self.name = name

    def _slice(self, X, X2):
        if isinstance(self.active_dims, slice):
            X = X[..., self.active_dims]
            if X2 is not None:
                X2 = X2[..., self.active_dims]
        else:
            X = tf.gather(X, self.active_dims, axis=-1)
            if X2 is not None:
                X2 = tf.gather(X2, self.active_dims, axis=-1)
        assert int(tf.shape(X)[-1]) == self.input_dim, "input_dim does not match sliced X"
        return X, X2


class Combination(Kernel):
    pass


class Sum(Combination):
    pass


class Product(Combination):
    pass
