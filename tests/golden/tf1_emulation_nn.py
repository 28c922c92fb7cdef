"""tf1_emulation + the TensorFlow-1.x ops of the reference's NETWORK code, so that the REFERENCE'S OWN SOURCE TEXT of the
Mixture-of-Experts (models/experts_n_est.py: the expert assignment and the calls of get_model :78-106, scale_manager_net,
conv_net_8g, conv_net_3g, normal_est_net, inception_module; utils/tf_util.py: conv3d, fully_connected, max_pool3d,
avg_pool3d, batch_norm_template, batch_norm_for_conv3d, batch_norm_for_fc, _variable_with_weight_decay,
_variable_on_cpu) can be executed in the build container, where TensorFlow 1.12 / Python 2.7 cannot be installed.

TEST INFRASTRUCTURE (fixture generation only; see tests/golden/make_golden.py::make_moe_tf_emulated).  Nothing here
restates the reference's architecture: layer widths, kernel sizes, scopes, the order of the concatenation, the channel
slices of the experts, the Python-2 integer division behind `np.round(128 / divider)` -- all of that runs from the
reference's text.  What IS restated is the documented semantics of the primitive ops:

  * tf.variable_scope / tf.get_variable / tf.Variable: variables live under '/'-joined scope names; their VALUES come from
    moe_weights.variable_value(name, shape) (a restored checkpoint, as test_n_est_w_experts.py:96 does), the initialisers
    are ignored;
  * tf.nn.conv3d (NDHWC, filter [kd, kh, kw, Cin, Cout], cross-correlation) and the pools with padding='SAME':
    out = ceil(n / stride), pad_total = max((out - 1) * stride + k - n, 0), the SMALLER half in front; max_pool3d ignores
    the padding cells, avg_pool3d divides by the number of VALID cells (TF's AvgPool does not count padding);
  * tf.nn.moments, tf.nn.batch_normalization(x, mean, var, beta, gamma, eps) = (x - mean) * gamma * rsqrt(var + eps) + beta;
  * tf.train.ExponentialMovingAverage.average(t): the shadow variable of the batch moment t (the moving statistics of a
    restored checkpoint); tf.cond on the python / numpy boolean fed for is_training; tf.no_op, control_dependencies,
    identity;
  * tf.nn.relu, tf.nn.softmax (last axis), tf.nn.bias_add, tf.matmul, tf.squeeze, tf.stack, tf.nn.l2_loss,
    tf.add_to_collection, tf.summary.scalar (no-ops).

Every op computes in float64 and rounds its RESULT to float32 (TensorFlow's kernels accumulate in float32 in an
unspecified order; the fixture is compared at 1e-4, three orders above that difference).
"""
import contextlib

import numpy as np

from tf1_emulation import *            # noqa: F401,F403
from tf1_emulation import Tensor, _t, _NN, Py2Int  # noqa: F401
import tf1_emulation as _base
import moe_weights

float16 = np.float16
bool = np.bool_                         # noqa: A001  (tf.bool)
uint16 = np.uint16

_scope_stack = []
created_variables = {}                  # name -> shape, in creation order (what a checkpoint of the graph would hold)


def reset_graph():
    del _scope_stack[:]
    created_variables.clear()


class _Scope(object):
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name):
    _scope_stack.append(str(name))
    try:
        yield _Scope("/".join(_scope_stack))
    finally:
        _scope_stack.pop()


@contextlib.contextmanager
def device(_name):
    yield


@contextlib.contextmanager
def control_dependencies(_ops):
    yield


def _scoped(name):
    return "/".join(_scope_stack + [name])


def _make_variable(name, shape):
    full = _scoped(name)
    shape = tuple(int(s) for s in shape)
    if full in created_variables:
        raise ValueError("Variable %s already exists" % full)      # tf.get_variable without reuse
    created_variables[full] = shape
    return Tensor(moe_weights.variable_value(full, shape))


def get_variable(name, shape, initializer=None, dtype=None):
    return _make_variable(name, shape)


def Variable(initial_value, name=None, trainable=True):            # noqa: N802
    return _make_variable(name, _t(initial_value).a.shape)


def constant(value, shape=None, dtype=None):                       # noqa: F811
    a = np.asarray(value, dtype=dtype)
    if shape is not None:
        a = np.broadcast_to(a, [int(s) for s in shape]).copy()
    return Tensor(a)


def constant_initializer(_v):
    return None


def truncated_normal_initializer(stddev=1.0):
    return None


def add_to_collection(_name, _value):
    return None


def no_op():
    return None


def identity(x):
    return x


def cond(pred, true_fn, false_fn):
    p = pred.a if isinstance(pred, Tensor) else pred
    return true_fn() if np.asarray(p).item() else false_fn()


def multiply(x, y, name=None):                                    # noqa: F811
    return _t(x) * y


def matmul(a, b):
    return Tensor((_t(a).a.astype(np.float64) @ _t(b).a.astype(np.float64)).astype(np.float32))


def squeeze(x):
    return Tensor(np.squeeze(_t(x).a))


def stack(values):
    return Tensor(np.stack([_t(v).a for v in values]))


def _same_pads(n, k, stride):
    out = -(-n // stride)
    total = max((out - 1) * stride + k - n, 0)
    return out, total // 2, total - total // 2


def _windows(x, ksize, strides, fill):
    """x [B, D, H, W, C] -> (padded array, output sizes, front pads) for a 'SAME' window op."""
    ks, st = [int(k) for k in ksize[1:4]], [int(s) for s in strides[1:4]]
    assert int(ksize[0]) == 1 and int(ksize[4]) == 1 and int(strides[0]) == 1 and int(strides[4]) == 1
    dims = x.shape[1:4]
    geo = [_same_pads(n, k, s) for n, k, s in zip(dims, ks, st)]
    pad = [(0, 0)] + [(g[1], g[2]) for g in geo] + [(0, 0)]
    return np.pad(x, pad, constant_values=fill), [g[0] for g in geo], ks, st


class _NNFull(_NN):
    @staticmethod
    def conv3d(inputs, kernel, strides, padding):
        assert padding == "SAME" and [int(s) for s in strides] == [1, 1, 1, 1, 1]
        x, f = _t(inputs).a.astype(np.float64), _t(kernel).a.astype(np.float64)
        kd, kh, kw, cin, cout = f.shape
        assert x.shape[4] == cin
        xp, out, _, _ = _windows(x, [1, kd, kh, kw, 1], [1, 1, 1, 1, 1], 0.0)
        y = np.zeros(x.shape[:1] + tuple(out) + (cout,), np.float64)
        for a in np.arange(kd):
            for b in np.arange(kh):
                for c in np.arange(kw):
                    y += xp[:, a:a + out[0], b:b + out[1], c:c + out[2], :] @ f[a, b, c]
        return Tensor(y.astype(np.float32))

    @staticmethod
    def bias_add(x, b):
        return Tensor(_t(x).a + _t(b).a)

    @staticmethod
    def relu(x):
        return Tensor(np.maximum(_t(x).a, np.float32(0)))

    @staticmethod
    def softmax(x):
        a = _t(x).a.astype(np.float64)
        e = np.exp(a - a.max(axis=-1, keepdims=True))
        return Tensor((e / e.sum(axis=-1, keepdims=True)).astype(np.float32))

    @staticmethod
    def l2_loss(x):
        return Tensor(np.float32(0.5 * np.sum(np.square(_t(x).a.astype(np.float64)))))

    @staticmethod
    def moments(x, axes, name=None):
        a = _t(x).a.astype(np.float64)
        axes = tuple(int(i) for i in axes)
        mean, var = Tensor(a.mean(axis=axes).astype(np.float32)), Tensor(a.var(axis=axes).astype(np.float32))
        mean.moment_of, var.moment_of = _scoped("moving_mean"), _scoped("moving_variance")
        return mean, var

    @staticmethod
    def batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
        a = _t(x).a.astype(np.float64)
        inv = _t(scale).a.astype(np.float64) / np.sqrt(_t(variance).a.astype(np.float64) + variance_epsilon)
        return Tensor((a * inv + (_t(offset).a.astype(np.float64) - _t(mean).a.astype(np.float64) * inv)).astype(np.float32))

    @staticmethod
    def max_pool3d(inputs, ksize, strides, padding, name=None):
        assert padding == "SAME"
        xp, out, ks, st = _windows(_t(inputs).a, ksize, strides, -np.inf)
        y = np.full(xp.shape[:1] + tuple(out) + xp.shape[4:], -np.inf, np.float32)
        for a in np.arange(ks[0]):
            for b in np.arange(ks[1]):
                for c in np.arange(ks[2]):
                    y = np.maximum(y, xp[:, a:a + st[0] * out[0]:st[0], b:b + st[1] * out[1]:st[1],
                                         c:c + st[2] * out[2]:st[2], :])
        return Tensor(y)

    @staticmethod
    def avg_pool3d(inputs, ksize, strides, padding, name=None):
        assert padding == "SAME"
        x = _t(inputs).a.astype(np.float64)
        xp, out, ks, st = _windows(x, ksize, strides, 0.0)
        ones, _, _, _ = _windows(np.ones(x.shape[:4] + (1,)), ksize, strides, 0.0)
        y = np.zeros(x.shape[:1] + tuple(out) + x.shape[4:], np.float64)
        cnt = np.zeros(x.shape[:1] + tuple(out) + (1,), np.float64)
        for a in np.arange(ks[0]):
            for b in np.arange(ks[1]):
                for c in np.arange(ks[2]):
                    sl = (slice(None), slice(a, a + st[0] * out[0], st[0]), slice(b, b + st[1] * out[1], st[1]),
                          slice(c, c + st[2] * out[2], st[2]), slice(None))
                    y += xp[sl]
                    cnt += ones[sl]
        return Tensor((y / cnt).astype(np.float32))


nn = _NNFull()


class _EMA(object):
    def __init__(self, decay):
        self.decay = decay

    def apply(self, _vars):
        raise RuntimeError("training branch taken: the fixture runs the inference graph (is_training = False)")

    def average(self, t):
        """Shadow variable of the batch moment `t` -- in a restored checkpoint, the moving statistic."""
        full = t.moment_of
        if full not in created_variables:
            created_variables[full] = t.a.shape
        return Tensor(moe_weights.variable_value(full, t.a.shape))


class _Train(object):
    ExponentialMovingAverage = _EMA


train = _Train()


class _Summary(object):
    @staticmethod
    def scalar(_name, _value):
        return None


summary = _Summary()


class _LayersNN(object):
    flatten = staticmethod(_base.contrib.layers.flatten)

    @staticmethod
    def xavier_initializer():
        return None


class _ContribNN(object):
    layers = _LayersNN()
    distributions = _base.contrib.distributions


contrib = _ContribNN()
