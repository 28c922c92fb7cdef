"""A numpy emulation of the ~25 TensorFlow-1.x ops that the reference's 3DmFV functions use, so that the REFERENCE'S OWN
SOURCE TEXT (utils/tf_util.py::get_3dmfv_n_est / get_3dmfv, the MuPS loop of models/experts_n_est.py::get_model) can be
executed in the build container, where TensorFlow 1.12 / Python 2.7 cannot be installed.

TEST INFRASTRUCTURE (fixture generation only; see tests/golden/make_golden.py::make_half2_tf_emulated).  Nothing here
restates the reference's algorithm: the statements that run are the reference's, extracted with `ast` from the files
under /root/reference and exec'd with `tf` bound to this module.  What IS restated is the documented semantics of the
primitive ops, each of which is a one-line numpy call:

  * dtypes as TF infers them: python scalars adopt the dtype of the tensor operand (float32 everywhere on this path),
    tf.range -> int32, comparisons -> bool;
  * tf.tile / expand_dims / concat / transpose / reshape / where / reduce_{sum,max,min} / square / exp / sqrt / abs /
    sign (sign(0) = 0) / pow / multiply / divide / cast / zeros_like / range, `x[:, tf.newaxis]` and `x[..., 0]` indexing;
  * tf.nn.l2_normalize(x, axis) = x * rsqrt(max(sum(x^2, axis, keepdims), 1e-12))  (TF 1.12 nn_impl.py);
  * tf.contrib.layers.flatten = reshape [B, -1];
  * tf.contrib.distributions.MultivariateNormalDiag(loc, scale_diag).prob(x) =
    exp(-0.5 sum_k ((x-loc)/scale)^2 - 0.5 D log(2 pi) - sum_k log(scale))  (TF 1.12 mvn_diag / normal log_prob);
  * `.shape[i].value` / `.get_shape()[i].value` return a Python-2 integer: `/` between two of them floors, as
    `n_points = points.get_shape()[1].value / n_rads` (experts_n_est.py:61) relies on.

Reductions run in float32 (numpy pairwise summation; Eigen's differs in association only).
"""
import numpy as np

newaxis = None
int32 = np.int32
float32 = np.float32


class Py2Int(int):
    """Python-2 integer: `/` floors."""

    def __truediv__(self, other):
        if isinstance(other, int):
            return Py2Int(int(self) // int(other))
        return int(self) / other

    def __rtruediv__(self, other):
        if isinstance(other, int):
            return Py2Int(int(other) // int(self))
        return other / int(self)

    def __mul__(self, other):
        r = int.__mul__(self, other)
        return Py2Int(r) if isinstance(other, int) and r is not NotImplemented else r

    __rmul__ = __mul__


class Dimension(object):
    def __init__(self, v):
        self.value = Py2Int(v)


def _arr(x, like=None):
    """Operand -> numpy array with the dtype TF would infer."""
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, np.ndarray):
        return x
    if like is not None:                       # python scalar adopts the tensor's dtype
        return np.asarray(x, dtype=like.dtype)
    if isinstance(x, float):
        return np.asarray(x, dtype=np.float32)
    if isinstance(x, (int, Py2Int)):
        return np.asarray(x, dtype=np.int32)
    return np.asarray(x)


class Tensor(object):
    __array_priority__ = 1000

    def __init__(self, a):
        a = np.asarray(a)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        if a.dtype == np.int64:
            a = a.astype(np.int32)
        self.a = a

    # ---- shape ----
    @property
    def shape(self):
        return [Dimension(v) for v in self.a.shape]

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self.a.dtype

    def __getitem__(self, item):
        return Tensor(self.a[item])

    # ---- arithmetic (python scalars adopt this tensor's dtype) ----
    def _bin(self, other, f, reverse=False):
        o = _arr(other, like=self.a)
        with np.errstate(all="ignore"):
            r = f(o, self.a) if reverse else f(self.a, o)
        return Tensor(r)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __gt__(self, o): return self._bin(o, np.greater)
    def __neg__(self): return Tensor(-self.a)


def _t(x):
    return x if isinstance(x, Tensor) else Tensor(_arr(x))


def constant(x, dtype=None):
    return Tensor(np.asarray(x, dtype=dtype))


def cast(x, dtype):
    return Tensor(_arr(x).astype(dtype))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_t(x).a, axis))


def tile(x, multiples):
    return Tensor(np.tile(_t(x).a, [int(m) for m in multiples]))


def concat(values, axis):
    return Tensor(np.concatenate([_t(v).a for v in values], axis=axis))


def transpose(x, perm=None):
    return Tensor(np.transpose(_t(x).a, perm))


def reshape(x, shape):
    return Tensor(np.reshape(_t(x).a, [int(s) for s in shape]))


def where(cond, x, y):
    return Tensor(np.where(_t(cond).a, _t(x).a, _t(y).a))


def zeros_like(x):
    return Tensor(np.zeros_like(_t(x).a))


def range(n):                                   # noqa: A001  (tf.range)
    return Tensor(np.arange(int(n), dtype=np.int32))


def reduce_sum(x, axis=None):
    a = _t(x).a
    return Tensor(np.sum(a, axis=axis, dtype=a.dtype))


def reduce_max(x, axis=None):
    return Tensor(np.max(_t(x).a, axis=axis))


def reduce_min(x, axis=None):
    return Tensor(np.min(_t(x).a, axis=axis))


def square(x):
    return Tensor(np.square(_t(x).a))


def exp(x):
    with np.errstate(all="ignore"):
        return Tensor(np.exp(_t(x).a))


def sqrt(x):
    with np.errstate(all="ignore"):
        return Tensor(np.sqrt(_t(x).a))


def abs(x):                                     # noqa: A001  (tf.abs)
    return Tensor(np.abs(_t(x).a))


def sign(x):
    return Tensor(np.sign(_t(x).a))


def pow(x, y):                                  # noqa: A001  (tf.pow)
    if isinstance(x, Tensor):
        a, b = x.a, _arr(y, like=x.a)
    elif isinstance(y, Tensor):
        b, a = y.a, _arr(x, like=y.a)
    else:                                       # two python scalars: float32 tensors
        a, b = np.asarray(x, np.float32), np.asarray(y, np.float32)
    with np.errstate(all="ignore"):
        return Tensor(np.power(a, b))


def multiply(x, y):
    return _t(x) * y


def divide(x, y):
    return _t(x) / y


class _NN(object):
    @staticmethod
    def l2_normalize(x, axis=None, epsilon=1e-12):
        a = _t(x).a
        sq = np.sum(np.square(a), axis=axis, keepdims=True, dtype=a.dtype)
        with np.errstate(all="ignore"):
            inv = (np.float32(1.0) / np.sqrt(np.maximum(sq, np.asarray(epsilon, a.dtype)))).astype(a.dtype)
        return Tensor(a * inv)


nn = _NN()


class _MVNDiag(object):
    def __init__(self, loc, scale_diag):
        self.loc, self.scale = _t(loc).a, _t(scale_diag).a

    def prob(self, x):
        x = _t(x).a
        with np.errstate(all="ignore"):
            z = ((x - self.loc) / self.scale).astype(np.float32)
            k = x.shape[-1]
            lp = (np.float32(-0.5) * np.sum(np.square(z), axis=-1, dtype=np.float32)
                  - np.float32(0.5 * k * np.log(2.0 * np.pi)) - np.sum(np.log(self.scale), axis=-1, dtype=np.float32))
            return Tensor(np.exp(lp.astype(np.float32)))


class _Layers(object):
    @staticmethod
    def flatten(x):
        a = _t(x).a
        return Tensor(a.reshape(a.shape[0], -1))


class _Distributions(object):
    MultivariateNormalDiag = _MVNDiag


class _Contrib(object):
    layers = _Layers()
    distributions = _Distributions()


contrib = _Contrib()
