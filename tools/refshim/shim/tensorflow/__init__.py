"""
NumPy-backed stand-in for the sliver of the TensorFlow 1.15 API that tgcsaba/GPSig's covariance path touches.

TEST INFRASTRUCTURE ONLY.  It exists so that `tests/golden/make_golden.py` can import the *unmodified* reference
sources from /root/reference (gpsig/signature_algs.py, kernels.py, inducing_variables.py, low_rank_calculations.py,
lags.py) in a container that has no TensorFlow, execute them eagerly in float64 and store their outputs as golden
vectors.  Nothing in the product (`gpsig_b200/`) imports this package.

Semantics implemented follow TF 1.15's documented behaviour for each op (eager evaluation instead of graph building).
`Tensor` is deliberately NOT an ndarray subclass: the reference builds `np.asarray([[M]])` object grids of tensors
(signature_algs.py:60) and NumPy must treat a tensor as an opaque scalar there, exactly like it does a tf.Tensor.
"""
import numpy as np

float64 = np.float64
float32 = np.float32
int32 = np.int32
int64 = np.int64

_rng = np.random.default_rng(0)


def set_shim_seed(seed):
    global _rng
    _rng = np.random.default_rng(seed)


class TShape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


def _a(x):
    """unwrap to ndarray / python scalar"""
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, Tensor) for e in x):
            return np.asarray([_a(e) for e in x])
    return x


def _idx(i):
    if isinstance(i, tuple):
        return tuple(_idx(e) for e in i)
    if isinstance(i, Tensor):
        return i.a
    if isinstance(i, slice):
        return slice(_ints(i.start), _ints(i.stop), _ints(i.step))
    return i


def _ints(v):
    if v is None:
        return None
    v = _a(v)
    return int(v)


class Tensor:
    __array_ufunc__ = None  # make numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, a):
        self.a = np.asarray(_a(a))

    # numpy must see an opaque scalar: sequence protocol present for slicing but len() is undefined (as in TF)
    def __len__(self):
        raise TypeError("len is not well defined for symbolic Tensors")

    def __iter__(self):
        raise TypeError("Tensor objects are only iterable when eager execution is enabled")

    @property
    def shape(self):
        return TShape(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    @property
    def ndim(self):
        return self.a.ndim

    def get_shape(self):
        return self.shape

    def __getitem__(self, i):
        return Tensor(self.a[_idx(i)])

    def __repr__(self):
        return "ShimTensor(%r)" % (self.a,)

    def __float__(self):
        return float(self.a)

    def __int__(self):
        return int(self.a)

    def __bool__(self):
        return bool(self.a)

    def __neg__(self):
        return Tensor(-self.a)

    def __add__(self, o): return Tensor(self.a + _a(o))
    def __radd__(self, o): return Tensor(_a(o) + self.a)
    def __sub__(self, o): return Tensor(self.a - _a(o))
    def __rsub__(self, o): return Tensor(_a(o) - self.a)
    def __mul__(self, o): return Tensor(self.a * _a(o))
    def __rmul__(self, o): return Tensor(_a(o) * self.a)
    def __truediv__(self, o): return Tensor(self.a / _a(o))
    def __rtruediv__(self, o): return Tensor(_a(o) / self.a)
    def __pow__(self, o): return Tensor(self.a ** _a(o))
    def __rpow__(self, o): return Tensor(_a(o) ** self.a)
    def __matmul__(self, o): return Tensor(self.a @ _a(o))
    def __lt__(self, o): return Tensor(self.a < _a(o))
    def __le__(self, o): return Tensor(self.a <= _a(o))
    def __gt__(self, o): return Tensor(self.a > _a(o))
    def __ge__(self, o): return Tensor(self.a >= _a(o))
    __hash__ = object.__hash__


def convert_to_tensor(x, dtype=None, **kw):
    return Tensor(np.asarray(_a(x), dtype=dtype))


def constant(x, dtype=None, **kw):
    return convert_to_tensor(x, dtype)


def shape(x):
    return np.asarray(np.shape(_a(x)), dtype=np.int64)


def unstack(x, axis=0):
    x = _a(x)
    if isinstance(x, np.ndarray) and x.ndim == 1 and np.issubdtype(x.dtype, np.integer):
        return [int(v) for v in x]
    return [Tensor(v) for v in np.moveaxis(np.asarray(x), axis, 0)]


def _shape_arg(s):
    s = _a(s)
    if np.ndim(s) == 0:
        return (int(s),)
    return tuple(int(_a(v)) for v in s)


def ones(shape, dtype=np.float64): return Tensor(np.ones(_shape_arg(shape), dtype=dtype))
def zeros(shape, dtype=np.float64): return Tensor(np.zeros(_shape_arg(shape), dtype=dtype))
def ones_like(x, dtype=None): return Tensor(np.ones_like(_a(x), dtype=dtype))
def zeros_like(x, dtype=None): return Tensor(np.zeros_like(_a(x), dtype=dtype))
def eye(n, dtype=np.float64): return Tensor(np.eye(int(_a(n)), dtype=dtype))
def fill(dims, value): return Tensor(np.full(_shape_arg(dims), _a(value)))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001 (mirrors tf.range)
    start = _a(start)
    if limit is None:
        r = np.arange(start, dtype=dtype)
    else:
        r = np.arange(start, _a(limit), _a(delta), dtype=dtype)
    return Tensor(r)


def cast(x, dtype): return Tensor(np.asarray(_a(x)).astype(dtype))
def reshape(x, shape): return Tensor(np.reshape(_a(x), _shape_arg(shape)))
def transpose(x, perm=None): return Tensor(np.transpose(_a(x), perm))
def expand_dims(x, axis): return Tensor(np.expand_dims(_a(x), axis))
def squeeze(x, axis=None): return Tensor(np.squeeze(_a(x), axis))
def tile(x, multiples): return Tensor(np.tile(_a(x), _shape_arg(multiples)))
def concat(values, axis): return Tensor(np.concatenate([np.asarray(_a(v)) for v in values], axis=axis))
def stack(values, axis=0): return Tensor(np.stack([np.asarray(_a(v)) for v in values], axis=axis))
def split(x, sizes, axis=0):
    x = _a(x)
    sizes = [int(_a(s)) for s in sizes]
    return [Tensor(p) for p in np.split(x, np.cumsum(sizes)[:-1], axis=axis)]
def reverse(x, axis): return Tensor(np.flip(_a(x), axis=tuple(axis)))


def add_n(inputs):
    out = np.asarray(_a(inputs[0])).copy()
    for t in inputs[1:]:
        out = out + _a(t)
    return Tensor(out)


def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


def reduce_sum(x, axis=None, keepdims=False): return Tensor(np.sum(_a(x), axis=_axis(axis), keepdims=keepdims))
def reduce_mean(x, axis=None, keepdims=False): return Tensor(np.mean(_a(x), axis=_axis(axis), keepdims=keepdims))
def reduce_prod(x, axis=None, keepdims=False): return Tensor(np.prod(_a(x), axis=_axis(axis), keepdims=keepdims))
def count_nonzero(x, axis=None): return Tensor(np.count_nonzero(_a(x), axis=_axis(axis)))
def argmax(x, axis=None): return Tensor(np.argmax(_a(x), axis=axis))


def cumsum(x, axis=0, exclusive=False, reverse=False):
    x = np.asarray(_a(x))
    if reverse:
        x = np.flip(x, axis)
    c = np.cumsum(x, axis=axis)
    if exclusive:
        # TF: exclusive cumsum [a,b,c] -> [0,a,a+b]; computed by shifting, not by subtracting
        c = np.roll(c, 1, axis=axis)
        sl = [slice(None)] * c.ndim
        sl[axis] = 0
        c[tuple(sl)] = 0
    if reverse:
        c = np.flip(c, axis)
    return Tensor(c)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = np.asarray(_a(a)), np.asarray(_a(b))
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return Tensor(np.matmul(a, b))


def exp(x): return Tensor(np.exp(_a(x)))
def log(x): return Tensor(np.log(_a(x)))
def sqrt(x): return Tensor(np.sqrt(_a(x)))
def square(x): return Tensor(np.square(_a(x)))
def cos(x): return Tensor(np.cos(_a(x)))
def sin(x): return Tensor(np.sin(_a(x)))
def floor(x): return Tensor(np.floor(_a(x)))
def ceil(x): return Tensor(np.ceil(_a(x)))
def abs(x): return Tensor(np.abs(_a(x)))  # noqa: A001
def maximum(x, y): return Tensor(np.maximum(_a(x), _a(y)))
def minimum(x, y): return Tensor(np.minimum(_a(x), _a(y)))
def where(c, x=None, y=None): return Tensor(np.where(_a(c), _a(x), _a(y)))
def gather(params, indices, axis=0): return Tensor(np.take(_a(params), np.asarray(_a(indices)), axis=axis))
def boolean_mask(x, mask, axis=0): return Tensor(np.compress(np.asarray(_a(mask)).astype(bool), _a(x), axis=axis))
def matrix_diag_part(x): return Tensor(np.diagonal(_a(x), axis1=-2, axis2=-1).copy())
def diag(x): return Tensor(np.diag(_a(x)))
def matrix_band_part(x, lo, hi):
    x = np.asarray(_a(x))
    n, m = x.shape[-2:]
    i, j = np.arange(n)[:, None], np.arange(m)[None, :]
    keep = np.ones((n, m), bool)
    if lo >= 0:
        keep &= (i - j) <= lo
    if hi >= 0:
        keep &= (j - i) <= hi
    return Tensor(x * keep)


def self_adjoint_eig(x):
    w, v = np.linalg.eigh(_a(x))
    return Tensor(w), Tensor(v)


def cholesky(x): return Tensor(np.linalg.cholesky(_a(x)))


# ---- randomness: deterministic numpy generators; every draw is also logged so that golden files can store it ----
draw_log = []


def _log(kind, val):
    draw_log.append((kind, np.asarray(val).copy()))
    return val


def random_uniform(shape, minval=0, maxval=None, dtype=np.float64, seed=None):
    if np.issubdtype(np.dtype(dtype), np.integer):
        return Tensor(_log("uniform_int", _rng.integers(minval, maxval, size=_shape_arg(shape)).astype(dtype)))
    hi = 1.0 if maxval is None else maxval
    return Tensor(_log("uniform", _rng.uniform(minval, hi, size=_shape_arg(shape)).astype(dtype)))


def random_normal(shape, dtype=np.float64, seed=None):
    return Tensor(_log("normal", _rng.standard_normal(_shape_arg(shape)).astype(dtype)))


def random_shuffle(x):
    x = np.asarray(_a(x))
    perm = _log("perm", _rng.permutation(x.shape[0]))
    return Tensor(x[perm])


class _NN:
    @staticmethod
    def top_k(x, k, sorted=True):  # noqa: A002
        x = np.asarray(_a(x))
        order = np.argsort(-x, kind="stable")[: int(k)]
        return Tensor(x[order]), Tensor(order)


nn = _NN()


class _Logging:
    ERROR = 40

    @staticmethod
    def set_verbosity(v):
        pass


logging = _Logging()


class Session:  # only so that reference modules that mention it import cleanly
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def run(self, x): return np.asarray(_a(x))
