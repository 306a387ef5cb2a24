"""A minimal TensorFlow-1.x graph-mode emulator over torch-CPU fp32 -- TEST INFRASTRUCTURE.

Purpose: the reference (wns349/tensorflow-yolo) builds its conv stack with TF1 calls
(net/layers.py:14,31-51,75-80,86,92-96,102,108,115,122,133) and TensorFlow cannot be
installed here.  This module provides *just* the TF API surface those call sites use, with
the documented TF semantics, so that the reference's own, unmodified builders
(net/v3.py:8-94, net/v2.py:10-60) and weight loader (net/base.py:26-46) can be executed
in the build container.  ``oracle/refimport.py`` installs it as ``sys.modules["tensorflow"]``.

It is NOT used by the product and is not a general TF replacement.  Semantics implemented
(each from the TF1 API docs, not from the reference):
  * tf.layers.conv2d: NHWC, HWIO kernel, 'SAME' (pad_total=(ceil(H/s)-1)*s+k-H, before=total//2)
    or 'VALID'; optional bias.
  * tf.layers.batch_normalization(training=False): (x-mean)*rsqrt(var+eps)*gamma+beta.
  * tf.nn.leaky_relu: max(alpha*x, x).
  * tf.layers.max_pooling2d 'VALID'; tf.pad CONSTANT zeros; tf.concat; tf.reshape; tf.identity.
  * tf.extract_image_patches(VALID, rates 1): output depth ordered (row, col, channel).
  * tf.image.resize_nearest_neighbor(align_corners=False): src = floor(dst*in/out).
"""
import contextlib
import math
import types

import numpy as np
import torch
import torch.nn.functional as F

float32 = "float32"


class _Shape(object):
    def __init__(self, dims):
        self._dims = list(dims)

    def as_list(self):
        return list(self._dims)


class Tensor(object):
    """Lazy graph node; ``fn(env)`` returns an NHWC torch.float32 tensor."""

    def __init__(self, shape, fn, name=None, op_type=None, inputs=()):
        self._shape = list(shape)
        self._fn = fn
        self.name = name
        self.op_type = op_type
        self.inputs = tuple(inputs)

    def get_shape(self):
        return _Shape(self._shape)

    @property
    def shape(self):
        return _Shape(self._shape)

    def __add__(self, other):
        assert self._shape == other._shape, (self._shape, other._shape)
        return Tensor(self._shape, lambda env: _eval(self, env) + _eval(other, env),
                      op_type="Add", inputs=(self, other))

    __hash__ = object.__hash__


def _eval(t, env):
    key = id(t)
    if key not in env:
        env[key] = t._fn(env)
    return env[key]


class _Op(object):
    def __init__(self, name):
        self.name = name


class Variable(object):
    def __init__(self, name, shape, init):
        self.op = _Op(name)
        self.name = name + ":0"
        self.shape = _Shape(shape)
        self.value = torch.as_tensor(np.asarray(init, dtype=np.float32).reshape(shape))


class _Graph(object):
    def __init__(self):
        self.variables = {}
        self.scopes = []

    def scoped(self, name):
        return "/".join(self.scopes + [name])

    def get_variable(self, name, shape, init):
        if name in self.variables:
            raise ValueError("Variable {} already exists".format(name))
        v = Variable(name, shape, init)
        self.variables[name] = v
        return v


_graph = _Graph()


def reset_default_graph():
    global _graph
    _graph = _Graph()


@contextlib.contextmanager
def variable_scope(scope):
    _graph.scopes.append(scope)
    try:
        yield
    finally:
        _graph.scopes.pop()


class GraphKeys(object):
    GLOBAL_VARIABLES = "variables"


def get_collection(key, scope=None):
    assert key == GraphKeys.GLOBAL_VARIABLES
    return [v for n, v in _graph.variables.items() if scope is None or n.startswith(scope)]


class _AssignOp(object):
    def __init__(self, var, value):
        self.var, self.value = var, value

    def run(self):
        self.var.value = torch.as_tensor(np.ascontiguousarray(self.value, dtype=np.float32))


def assign(var, value, validate_shape=True):
    value = np.asarray(value)
    if validate_shape and list(value.shape) != var.shape.as_list():
        raise ValueError("shape mismatch assigning {}: {} vs {}".format(
            var.name, value.shape, var.shape.as_list()))
    return _AssignOp(var, value)


def placeholder(dtype, shape, name=None):
    t = Tensor(shape, None, name=name, op_type="Placeholder")
    t._fn = lambda env: env[("feed", id(t))]
    return t


def pad(t, paddings, mode="CONSTANT"):
    assert mode == "CONSTANT"
    (n0, n1), (h0, h1), (w0, w1), (c0, c1) = paddings
    assert (n0, n1, c0, c1) == (0, 0, 0, 0)
    s = t._shape
    out_shape = [s[0], s[1] + h0 + h1, s[2] + w0 + w1, s[3]]
    return Tensor(out_shape, lambda env: F.pad(_eval(t, env), (0, 0, w0, w1, h0, h1)),
                  op_type="Pad", inputs=(t,))


def _same_pads(size, k, s):
    out = int(math.ceil(size / s))
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _conv2d(inputs, filters, kernel_size, padding="valid", strides=(1, 1), use_bias=True, name=None):
    k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
    s = strides if isinstance(strides, int) else strides[0]
    n, h, w, cin = inputs._shape
    limit = math.sqrt(6.0 / (k * k * cin + k * k * filters))  # TF1 default: glorot_uniform
    rng = np.random.RandomState(len(_graph.variables))
    kernel = _graph.get_variable(_graph.scoped(name + "/kernel"), [k, k, cin, filters],
                                 rng.uniform(-limit, limit, size=(k, k, cin, filters)))
    bias = _graph.get_variable(_graph.scoped(name + "/bias"), [filters],
                               np.zeros(filters)) if use_bias else None
    if padding.upper() == "SAME":
        ph, pw = _same_pads(h, k, s), _same_pads(w, k, s)
        ho, wo = int(math.ceil(h / s)), int(math.ceil(w / s))
    else:
        ph = pw = (0, 0)
        ho, wo = (h - k) // s + 1, (w - k) // s + 1

    def fn(env):
        x = _eval(inputs, env).permute(0, 3, 1, 2)
        x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
        wt = kernel.value.permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
        y = F.conv2d(x, wt, None if bias is None else bias.value, stride=s)
        return y.permute(0, 2, 3, 1).contiguous()

    return Tensor([n, ho, wo, filters], fn, name=name, op_type="Conv2D", inputs=(inputs,))


def _batch_normalization(x, training=False, momentum=0.99, epsilon=1e-3, name=None):
    assert not training, "the oracle only covers inference (TEST mode)"
    c = x._shape[-1]
    beta = _graph.get_variable(_graph.scoped(name + "/beta"), [c], np.zeros(c))
    gamma = _graph.get_variable(_graph.scoped(name + "/gamma"), [c], np.ones(c))
    mean = _graph.get_variable(_graph.scoped(name + "/moving_mean"), [c], np.zeros(c))
    var = _graph.get_variable(_graph.scoped(name + "/moving_variance"), [c], np.ones(c))

    def fn(env):
        v = _eval(x, env)
        inv = torch.rsqrt(var.value + epsilon) * gamma.value
        return v * inv + (beta.value - mean.value * inv)

    return Tensor(x._shape, fn, name=name, op_type="FusedBatchNorm", inputs=(x,))


def _max_pooling2d(inputs, pool_size, strides, padding="valid"):
    assert padding.upper() == "VALID"
    n, h, w, c = inputs._shape
    ho, wo = (h - pool_size) // strides + 1, (w - pool_size) // strides + 1

    def fn(env):
        x = _eval(inputs, env).permute(0, 3, 1, 2)
        return F.max_pool2d(x, pool_size, strides).permute(0, 2, 3, 1).contiguous()

    return Tensor([n, ho, wo, c], fn, op_type="MaxPool", inputs=(inputs,))


def _leaky_relu(x, alpha=0.2, name=None):
    return Tensor(x._shape, lambda env: torch.maximum(_eval(x, env) * alpha, _eval(x, env)),
                  name=name, op_type="LeakyRelu", inputs=(x,))


def concat(values, axis):
    shp = list(values[0]._shape)
    shp[axis] = sum(v._shape[axis] for v in values)
    return Tensor(shp, lambda env: torch.cat([_eval(v, env) for v in values], dim=axis),
                  op_type="ConcatV2", inputs=tuple(values))


def extract_image_patches(images, ksizes, strides, rates, padding):
    assert padding == "VALID" and list(rates) == [1, 1, 1, 1]
    kh, kw = ksizes[1], ksizes[2]
    sh, sw = strides[1], strides[2]
    assert (kh, kw) == (sh, sw)
    n, h, w, c = images._shape

    def fn(env):
        x = _eval(images, env)
        b = x.shape[0]
        x = x.reshape(b, h // kh, kh, w // kw, kw, c).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(b, h // kh, w // kw, kh * kw * c).contiguous()

    return Tensor([n, h // kh, w // kw, kh * kw * c], fn, op_type="ExtractImagePatches", inputs=(images,))


def _resize_nearest_neighbor(images, size):
    n, h, w, c = images._shape
    oh, ow = size

    def fn(env):
        x = _eval(images, env)
        yi = torch.floor(torch.arange(oh) * (h / oh)).long().clamp(max=h - 1)
        xi = torch.floor(torch.arange(ow) * (w / ow)).long().clamp(max=w - 1)
        return x[:, yi][:, :, xi].contiguous()

    return Tensor([n, oh, ow, c], fn, op_type="ResizeNearestNeighbor", inputs=(images,))


def reshape(t, shape):
    known = [d for d in t._shape if d is not None]
    total = int(np.prod(known))
    shp = list(shape)
    if t._shape[0] is None:
        assert shp[0] == -1
        out_shape = [None] + shp[1:]
        assert int(np.prod(shp[1:])) == total
    else:
        out_shape = shp
    return Tensor(out_shape, lambda env: _eval(t, env).reshape(shp), op_type="Reshape", inputs=(t,))


def identity(t, name=None):
    return Tensor(t._shape, lambda env: _eval(t, env), name=name, op_type="Identity", inputs=(t,))


layers = types.SimpleNamespace(conv2d=_conv2d, batch_normalization=_batch_normalization,
                               max_pooling2d=_max_pooling2d)
nn = types.SimpleNamespace(leaky_relu=_leaky_relu)
image = types.SimpleNamespace(resize_nearest_neighbor=_resize_nearest_neighbor)


def run(fetch, feed_dict=None):
    """Stand-in for ``sess.run``: evaluates a Tensor (fp32 numpy out) or a list of assign ops."""
    if isinstance(fetch, (list, tuple)) and all(isinstance(f, _AssignOp) for f in fetch):
        for f in fetch:
            f.run()
        return None
    env = {}
    for ph, val in (feed_dict or {}).items():
        env[("feed", id(ph))] = torch.as_tensor(np.asarray(val, dtype=np.float32))
    with torch.no_grad():
        return _eval(fetch, env).numpy()
