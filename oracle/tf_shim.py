"""A numpy stand-in for the handful of TensorFlow primitives that the reference's components/loss.py and
components/matting_v2.py call.  TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py in the authoring container).

Purpose: TensorFlow cannot be installed here, but the reference's OWN composition logic -- integral images and their
slicing, the [r+1, r] paddings, the 1/len(args) of iter_on_layers, the normalisers of the style term, the dict order of
the total -- can be executed if the primitives underneath are provided.  Each primitive below is unambiguous
(reshape, matmul, cumsum, inverse, mean ...) except two, which follow TensorFlow's documented behaviour:
  * tf.pad(mode='SYMMETRIC')  == numpy.pad(mode='symmetric')  (the border value is repeated);
  * tf.image.resize(images, size) in TF2: bilinear, half-pixel centres, no antialiasing:
        src = (i + 0.5) * in/out - 0.5;  lower = max(floor(src), 0);  upper = min(ceil(src), in - 1);  w = src - floor(src).
Golden vectors produced through this shim therefore pin the RESTATEMENT OF THE REFERENCE'S CODE, not TensorFlow's kernels.
"""
import types

import numpy as np


class _LinearOperator:
    """tf.linalg.LinearOperator as far as matting_v2.py / loss.py use it: .matmul -> _matmul, .shape -> _shape()."""

    def __init__(self, dtype, graph_parents=None, is_self_adjoint=None, is_positive_definite=None, name=None):
        self.dtype = dtype

    def matmul(self, x, adjoint=False, adjoint_arg=False):
        return self._matmul(x, adjoint=adjoint, adjoint_arg=adjoint_arg)

    @property
    def shape(self):
        return self._shape()


def _resize_axis(a, out, axis):
    n = a.shape[axis]
    if out == n:
        return a
    i = np.arange(out, dtype=np.float64)
    src = (i + 0.5) * (float(n) / float(out)) - 0.5
    f = np.floor(src)
    lo = np.maximum(f, 0).astype(np.int64)
    hi = np.minimum(np.ceil(src), n - 1).astype(np.int64)
    w = (src - f).astype(a.dtype)
    shape = [1] * a.ndim
    shape[axis] = out
    w = w.reshape(shape)
    return np.take(a, lo, axis=axis) + (np.take(a, hi, axis=axis) - np.take(a, lo, axis=axis)) * w


def _resize(images, size):
    a = np.asarray(images)
    return _resize_axis(_resize_axis(a, int(size[0]), 1), int(size[1]), 2)


def _cast(x, dtype=None):
    return np.asarray(x).astype(dtype)


def _constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(value, dtype=dtype)
    return np.full(tuple(shape), a, dtype=a.dtype) if shape is not None else a


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a = np.swapaxes(a, -1, -2) if transpose_a else a
    b = np.swapaxes(b, -1, -2) if transpose_b else b
    return a @ b


def _eye(n, batch_shape=None, dtype=np.float32):
    e = np.eye(n, dtype=dtype)
    return np.broadcast_to(e, tuple(batch_shape) + (n, n)).copy() if batch_shape is not None else e


def _pad(t, paddings, mode="CONSTANT"):
    return np.pad(t, np.asarray(paddings), mode={"SYMMETRIC": "symmetric", "CONSTANT": "constant"}[mode])


def make_module():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32 = np.float32, np.float64, np.int32
    tf.linalg = types.SimpleNamespace(LinearOperator=_LinearOperator, inv=np.linalg.inv)
    tf.math = types.SimpleNamespace(squared_difference=lambda a, b: (np.asarray(a) - np.asarray(b)) ** 2)
    tf.image = types.SimpleNamespace(resize=_resize)
    tf.TensorShape = tuple
    tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)
    tf.constant = _constant
    tf.cast = _cast
    tf.identity = lambda x, name=None: np.array(x)
    tf.expand_dims = lambda a, axis: np.expand_dims(a, axis)
    tf.reshape = lambda t, shape, name=None: np.reshape(t, shape)
    tf.transpose = lambda t, perm=None: np.transpose(t, perm)
    tf.cumsum = lambda t, axis=0: np.cumsum(t, axis=axis)
    tf.pad = _pad
    tf.eye = _eye
    tf.matmul = _matmul
    tf.add_n = lambda xs: sum(xs[1:], xs[0])
    tf.reduce_mean = lambda input_tensor, axis=None: np.mean(input_tensor, axis=axis)
    tf.reduce_sum = lambda input_tensor, axis=None, name=None: np.sum(input_tensor, axis=axis)
    tf.square = np.square
    tf.squeeze = np.squeeze
    tf.multiply = np.multiply
    tf.range = lambda a, b, dtype=None: np.arange(a, b, dtype=dtype)
    tf.ones = lambda shape, dtype=np.float32: np.ones(shape, dtype=dtype)
    return tf
