"""ORACLE — TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py).

A NumPy stand-in for the handful of TensorFlow ops that the reference's
``tf_fast_rnnt/python/tf_fast_rnnt/rnnt_loss.py`` calls, so that file can be
imported and executed *unmodified* in a container without TensorFlow and its
outputs frozen as golden vectors.  Only the semantics the reference relies on
are implemented (eager NumPy arrays, float32 default for Python floats, int32
true division -> float64, first-index argmax, sequential cumsum).

``install()`` registers the stand-in as ``sys.modules['tensorflow']`` and a
stub ``tf_fast_rnnt`` module whose two custom ops (the lattice recursion and
cummin — compiled CUDA in the reference, tf_fast_rnnt_op.cc:27-38) come from
the plain-C oracle; ``load_reference()`` then imports the reference file from
/root/reference by path.  Nothing here is reachable from the product.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np


class _DType:
    def __init__(self, np_type):
        self.np = np_type
        if np.issubdtype(np_type, np.integer):
            self.max = int(np.iinfo(np_type).max)
            self.min = int(np.iinfo(np_type).min)

    def __repr__(self):
        return f"tf_emu.{np.dtype(self.np).name}"


class ETensor(np.ndarray):
    """ndarray whose augmented assignments rebind instead of mutating — TF
    tensors are immutable, and the reference writes ``py -= normalizers`` on
    what NumPy would treat as a view of its input (rnnt_loss.py:445)."""

    def __iadd__(self, o):
        return np.add(self, o)

    def __isub__(self, o):
        return np.subtract(self, o)

    def __imul__(self, o):
        return np.multiply(self, o)

    def __itruediv__(self, o):
        return np.true_divide(self, o)


def _t(a):
    return np.asarray(a).view(ETensor)


def _npdt(d):
    if d is None:
        return None
    if isinstance(d, _DType):
        return d.np
    return np.dtype(d).type


def _arr(x, dtype=None):
    """Python floats become float32, Python ints int32 (TF's defaults)."""
    if isinstance(x, np.ndarray) or isinstance(x, np.generic):
        a = np.asarray(x)
    elif isinstance(x, (list, tuple)) and any(
            isinstance(v, (np.ndarray, np.generic)) for v in x):
        a = np.asarray(x)
    else:
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    if dtype is not None:
        a = a.astype(_npdt(dtype))
    return a


def _shape_list(s):
    return tuple(int(v) for v in np.asarray(s).reshape(-1))


def make_module():
    tf = types.ModuleType("tensorflow")
    tf.__emulated__ = True
    tf.float32 = _DType(np.float32)
    tf.float64 = _DType(np.float64)
    tf.int32 = _DType(np.int32)
    tf.int64 = _DType(np.int64)
    tf.newaxis = None
    tf.Tensor = np.ndarray
    tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)

    tf.shape = lambda x: np.asarray(np.shape(x), dtype=np.int32)
    tf.convert_to_tensor = lambda x, dtype=None: _t(_arr(x, dtype))
    tf.cast = lambda x, dtype: np.asarray(x).astype(_npdt(dtype))
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), _shape_list(shape)))
    tf.transpose = lambda x, perm=None: _t(np.transpose(np.asarray(x), perm))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(np.asarray(x), axis))
    tf.squeeze = lambda x, axis=None: _t(np.squeeze(np.asarray(x), axis))
    tf.concat = lambda values, axis: _t(np.concatenate([_arr(v) for v in values], axis=axis))
    tf.stack = lambda values, axis=0: np.stack([_arr(v) for v in values], axis=axis)
    tf.tile = lambda x, multiples: np.tile(_arr(x), _shape_list(multiples))
    tf.reverse = lambda x, axis: np.flip(np.asarray(x), axis=tuple(axis))
    tf.where = lambda c, a, b: np.where(c, a, b)
    tf.sigmoid = lambda x: (1.0 / (1.0 + np.exp(-np.asarray(x)))).astype(np.asarray(x).dtype)
    tf.matmul = lambda a, b, transpose_b=False: np.matmul(
        a, np.swapaxes(b, -1, -2) if transpose_b else b)

    def fill(dims, value):
        return np.full(_shape_list(dims), _arr(value))
    tf.fill = fill

    def zeros(shape, dtype=tf.float32):
        return np.zeros(_shape_list(shape), dtype=_npdt(dtype))
    tf.zeros = zeros

    def broadcast_to(x, shape):
        return np.broadcast_to(_arr(x), _shape_list(shape)).copy()
    tf.broadcast_to = broadcast_to

    def range_(start, limit=None, delta=1, dtype=None):
        if limit is None:
            start, limit = 0, start
        a = np.arange(int(start), int(limit), int(delta))
        return a.astype(_npdt(dtype) if dtype is not None else np.int32)
    tf.range = range_

    def meshgrid(*args, indexing="xy"):
        return np.meshgrid(*args, indexing=indexing)
    tf.meshgrid = meshgrid

    def clip_by_value(x, clip_value_min, clip_value_max):
        return np.clip(x, clip_value_min, clip_value_max).astype(np.asarray(x).dtype)
    tf.clip_by_value = clip_by_value

    def gather_nd(params, indices, batch_dims=0):
        params = np.asarray(params)
        indices = np.asarray(indices)
        q = indices.shape[-1]
        idx = []
        for d in range(batch_dims):
            shp = [1] * (indices.ndim - 1)
            shp[d] = params.shape[d]
            idx.append(np.arange(params.shape[d]).reshape(shp))
        for k in range(q):
            idx.append(indices[..., k])
        return _t(params[tuple(idx)])
    tf.gather_nd = gather_nd

    def gather(params, indices, batch_dims=0, axis=None):
        params = np.asarray(params)
        indices = np.asarray(indices)
        if axis is None:
            axis = batch_dims
        if batch_dims == 0:
            return np.take(params, indices, axis=axis)
        assert batch_dims == 1 and axis == 1
        b = np.arange(params.shape[0]).reshape((-1,) + (1,) * (indices.ndim - 1))
        return params[b, indices]
    tf.gather = gather

    def tensor_scatter_nd_update(tensor, indices, updates):
        out = np.array(tensor, copy=True)
        indices = np.asarray(indices)
        out[tuple(indices[..., k] for k in range(indices.shape[-1]))] = updates
        return out
    tf.tensor_scatter_nd_update = tensor_scatter_nd_update

    def cumsum(x, axis=0):
        x = np.asarray(x)
        return np.cumsum(x, axis=axis, dtype=x.dtype)
    tf.cumsum = cumsum

    def reduce_sum(x, axis=None, keepdims=False):
        x = np.asarray(x)
        return np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype)

    def reduce_mean(x, axis=None, keepdims=False):
        x = np.asarray(x)
        return np.mean(x, axis=axis, keepdims=keepdims, dtype=x.dtype)

    def reduce_max(x, axis=None, keepdims=False):
        return np.max(np.asarray(x), axis=axis, keepdims=keepdims)

    def reduce_logsumexp(x, axis=None, keepdims=False):
        x = np.asarray(x)
        m = np.max(x, axis=axis, keepdims=True)
        m = np.where(np.isfinite(m), m, 0).astype(x.dtype)
        r = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True, dtype=x.dtype)) + m
        return r if keepdims else np.squeeze(r, axis=axis)

    def argmax(x, axis=None, output_type=tf.int64):
        return np.argmax(np.asarray(x), axis=axis).astype(_npdt(output_type))

    tf.reduce_sum, tf.reduce_mean, tf.reduce_max = reduce_sum, reduce_mean, reduce_max
    tf.argmax = argmax

    m = types.ModuleType("tensorflow.math")
    m.reduce_max, m.reduce_sum, m.reduce_mean = reduce_max, reduce_sum, reduce_mean
    m.reduce_logsumexp = reduce_logsumexp
    m.argmax = argmax
    m.exp = lambda x: np.exp(np.asarray(x))
    m.log = lambda x: np.log(np.asarray(x))
    m.nextafter = lambda a, b: np.nextafter(np.float32(a), np.float32(b))
    tf.math = m

    la = types.ModuleType("tensorflow.linalg")
    la.matvec = lambda a, b: np.matmul(np.asarray(a), np.asarray(b))
    tf.linalg = la
    return tf


REFERENCE_RNNT_LOSS = (
    "/root/reference/tf_fast_rnnt/python/tf_fast_rnnt/rnnt_loss.py")


def install():
    """Register the stand-in ``tensorflow`` and a stub ``tf_fast_rnnt``."""
    from . import rnnt_oracle as orc

    tf = make_module()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.math"] = tf.math
    sys.modules["tensorflow.linalg"] = tf.linalg

    stub = types.ModuleType("tf_fast_rnnt")

    def mutual_information_recursion(px, py, boundary, calc_gradients=False):
        # the custom op "FastRNNTLoss" (tf_fast_rnnt_op.cc:27-34), float32
        return orc.mutual_information_recursion(
            np.asarray(px, np.float32), np.asarray(py, np.float32),
            np.asarray(boundary, np.int32), calc_gradients, np.float32)

    stub.mutual_information_recursion = mutual_information_recursion
    stub.cummin = lambda x: orc.cummin(np.ascontiguousarray(x, dtype=np.int32))
    sys.modules["tf_fast_rnnt"] = stub
    return tf


def load_reference(path: str = REFERENCE_RNNT_LOSS):
    """Import the reference's rnnt_loss.py (read-only, by path)."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    install()
    old = np.get_printoptions()
    spec = importlib.util.spec_from_file_location("_reference_rnnt_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    np.set_printoptions(**old)   # the reference changes print options at import
    return mod
