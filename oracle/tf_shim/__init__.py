"""Tier-A oracle support: an eager, NumPy-backed stand-in for the `tensorflow` 1.x
symbols that the reference's box-level hot path touches.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is imported by the product
package; only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s CPU-baseline leg
and `oracle/gen_golden.py` use it.

Why it exists: the reference (`/root/reference`) is TensorFlow-1.x graph code and
TensorFlow cannot be installed in this image.  With this shim on `sys.meta_path`
the reference files

    utils/net_tools.py  utils/common_tools.py
    utils/tf_extended/{bboxes,tensors,math}.py

are imported and EXECUTED UNMODIFIED; every `tf.*` primitive they call is
evaluated immediately with NumPy in float32, one op at a time (which is also how
TF's Eigen CPU kernels evaluate: no fusion, no FMA contraction).

Third-party arithmetic that is NOT in /root/reference and is restated here from
TensorFlow's documented behaviour (TensorFlow 1.x, version unpinned by the
reference, API usage implies 1.12-1.15):
  * `tf.nn.top_k`       descending, equal values -> lower index first.
  * `tf.argmax/argmin`  first occurrence.
  * `tf.image.non_max_suppression` (NonMaxSuppressionV3, score_threshold=-inf):
      greedy over candidates in descending score (ties -> lower index; TF<=1.14
      leaves heap tie order unspecified, BASELINE.json fixes lowest index),
      suppress iff IoU > iou_threshold, IoU as in
      tensorflow/core/kernels/non_max_suppression_op.cc (min/max-normalised
      corners, 0 when either area <= 0, inter / (area_i + area_j - inter)).
  * `tf.exp/tf.log`     float32 results; NumPy's float32 libm is used, the real
      Eigen kernels may differ in the last ulp (tolerance-only parity, 1e-5).
Those four are "parity unpinned" by the reference itself (it has no tests).
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

import numpy as np

__all__ = ["install", "Tensor", "to_numpy"]


# --------------------------------------------------------------------------- #
# Tensor
# --------------------------------------------------------------------------- #
class _Shape:
    def __init__(self, dims):
        self._dims = [None if d is None else int(d) for d in dims]

    def as_list(self):
        return list(self._dims)

    def is_fully_defined(self):
        return all(d is not None for d in self._dims)

    def with_rank(self, rank):
        assert len(self._dims) == rank
        return self

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __getitem__(self, i):
        return self._dims[i]


def _unwrap(x):
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(e) for e in x)
    return x


def to_numpy(x):
    """Recursively convert Tensors (in lists / tuples / dicts) to ndarrays."""
    if isinstance(x, Tensor):
        return np.asarray(x.a)
    if isinstance(x, dict):
        return {k: to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_numpy(e) for e in x)
    return x


def _idx(i):
    if isinstance(i, Tensor):
        a = i.a
        return int(a) if np.ndim(a) == 0 else np.asarray(a)
    if isinstance(i, tuple):
        return tuple(_idx(e) for e in i)
    return i


class Tensor:
    """Eager tensor: a thin wrapper over an ndarray with TF-1 style accessors."""

    __array_ufunc__ = None  # make ndarray <op> Tensor defer to Tensor.__r<op>__
    __slots__ = ("a",)

    def __init__(self, a):
        self.a = np.asarray(_unwrap(a))

    # -- TF-style metadata
    def get_shape(self):
        return _Shape(self.a.shape)

    @property
    def shape(self):
        return _Shape(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def numpy(self):
        return self.a

    def __repr__(self):
        return "shim.Tensor(%r)" % (self.a,)

    def __len__(self):
        return len(self.a)

    def __iter__(self):
        for i in range(self.a.shape[0]):
            yield Tensor(self.a[i])

    def __bool__(self):
        return bool(self.a)

    def __int__(self):
        return int(self.a)

    def __index__(self):
        return int(self.a)

    def __float__(self):
        return float(self.a)

    def __getitem__(self, i):
        return Tensor(self.a[_idx(i)])

    # -- arithmetic: one NumPy op per TF op, operand dtype preserved (float32
    #    tensors with Python scalars stay float32 under NumPy>=2 promotion rules,
    #    which is what TF's convert_to_tensor(dtype=other.dtype) does).
    def _bin(self, other, fn, rev=False):
        o = _unwrap(other)
        with np.errstate(all="ignore"):
            return Tensor(fn(o, self.a) if rev else fn(self.a, o))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    def __neg__(self): return Tensor(np.negative(self.a))
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    __hash__ = object.__hash__


def _t(x):
    return x if isinstance(x, Tensor) else Tensor(x)


def _shape_arg(shape):
    s = _unwrap(shape)
    if isinstance(s, np.ndarray):
        s = s.tolist()
    if isinstance(s, _Shape):
        s = s.as_list()
    return tuple(int(_unwrap(d)) for d in s)


def _dt(dtype):
    return None if dtype is None else np.dtype(dtype)


# --------------------------------------------------------------------------- #
# The `tensorflow` module surface (only what the hot path uses)
# --------------------------------------------------------------------------- #
def _build_tf_module():
    tf = types.ModuleType("tensorflow")
    tf.__path__ = []  # behave as a package so submodule imports are attempted

    tf.float32 = np.dtype("float32")
    tf.float64 = np.dtype("float64")
    tf.int32 = np.dtype("int32")
    tf.int64 = np.dtype("int64")
    tf.bool = np.dtype("bool")
    tf.Tensor = Tensor

    def TensorShape(dims):
        return _Shape(dims)
    tf.TensorShape = TensorShape

    @contextlib.contextmanager
    def name_scope(name=None, default_name=None, values=None):
        yield name or default_name
    tf.name_scope = name_scope

    @contextlib.contextmanager
    def device(_):
        yield
    tf.device = device

    def constant(value, dtype=None, shape=None, name=None):
        a = np.asarray(_unwrap(value), dtype=_dt(dtype))
        if dtype is None and a.dtype == np.float64:
            a = a.astype(np.float32)       # TF default float is float32
        if dtype is None and a.dtype == np.int64:
            a = a.astype(np.int32)         # TF default int is int32
        if shape is not None:
            a = np.broadcast_to(a, _shape_arg(shape)).copy()
        return Tensor(a)
    tf.constant = constant
    tf.convert_to_tensor = lambda v, dtype=None, name=None: (
        Tensor(np.asarray(_unwrap(v), dtype=_dt(dtype))))

    def _un(fn):
        def op(x, name=None):
            with np.errstate(all="ignore"):
                return Tensor(fn(_t(x).a))
        return op

    def _bi(fn):
        def op(x, y, name=None):
            xa, ya = _unwrap(x), _unwrap(y)
            with np.errstate(all="ignore"):
                return Tensor(fn(xa, ya))
        return op

    tf.exp = _un(np.exp)
    tf.log = _un(np.log)
    tf.abs = _un(np.abs)
    tf.sqrt = _un(np.sqrt)
    tf.zeros_like = _un(np.zeros_like)
    tf.ones_like = _un(np.ones_like)
    tf.logical_not = _un(np.logical_not)
    tf.maximum = _bi(np.maximum)
    tf.minimum = _bi(np.minimum)
    tf.equal = _bi(np.equal)
    tf.not_equal = _bi(np.not_equal)
    tf.less = _bi(np.less)
    tf.greater = _bi(np.greater)
    tf.greater_equal = _bi(np.greater_equal)
    tf.less_equal = _bi(np.less_equal)
    tf.logical_and = _bi(np.logical_and)
    tf.logical_or = _bi(np.logical_or)
    tf.divide = _bi(np.true_divide)
    tf.multiply = _bi(np.multiply)
    tf.add = _bi(np.add)
    tf.subtract = _bi(np.subtract)

    def cast(x, dtype, name=None):
        return Tensor(_t(x).a.astype(np.dtype(dtype)))
    tf.cast = cast

    def shape(x, name=None, out_type=None):
        return Tensor(np.asarray(np.shape(_unwrap(x)), dtype=np.int32))
    tf.shape = shape

    def size(x, name=None):
        return Tensor(np.asarray(np.size(_unwrap(x)), dtype=np.int32))
    tf.size = size

    def reshape(x, shape, name=None):
        return Tensor(np.reshape(_t(x).a, _shape_arg(shape)))
    tf.reshape = reshape

    def stack(values, axis=0, name=None):
        return Tensor(np.stack([np.asarray(_unwrap(v)) for v in values], axis=axis))
    tf.stack = stack

    def unstack(value, num=None, axis=0, name=None):
        a = _t(value).a
        return [Tensor(np.take(a, i, axis=axis)) for i in range(a.shape[axis])]
    tf.unstack = unstack

    def concat(values, axis, name=None):
        return Tensor(np.concatenate([np.asarray(_unwrap(v)) for v in values], axis=axis))
    tf.concat = concat

    def expand_dims(x, axis=None, name=None, dim=None):
        return Tensor(np.expand_dims(_t(x).a, axis if axis is not None else dim))
    tf.expand_dims = expand_dims

    def transpose(x, perm=None, name=None):
        return Tensor(np.transpose(_t(x).a, perm))
    tf.transpose = transpose

    def zeros(shape, dtype=tf.float32, name=None):
        return Tensor(np.zeros(_shape_arg(shape), dtype=np.dtype(dtype)))
    tf.zeros = zeros

    def ones(shape, dtype=tf.float32, name=None):
        return Tensor(np.ones(_shape_arg(shape), dtype=np.dtype(dtype)))
    tf.ones = ones

    def range_(start, limit=None, delta=1, dtype=None, name=None):
        lo, hi = (0, _unwrap(start)) if limit is None else (_unwrap(start), _unwrap(limit))
        return Tensor(np.arange(lo, hi, _unwrap(delta), dtype=_dt(dtype) or np.int32))
    tf.range = range_

    def _red(fn):
        def op(x, axis=None, keepdims=False, name=None, keep_dims=None,
               reduction_indices=None):
            if keep_dims is not None:
                keepdims = keep_dims
            ax = axis if axis is not None else reduction_indices
            if isinstance(ax, list):
                ax = tuple(ax)
            return Tensor(fn(_t(x).a, axis=ax, keepdims=keepdims))
        return op
    tf.reduce_sum = _red(np.sum)
    tf.reduce_max = _red(np.max)
    tf.reduce_min = _red(np.min)

    def count_nonzero(x, axis=None, dtype=tf.int64, name=None):
        return Tensor(np.asarray(np.count_nonzero(_t(x).a, axis=axis), dtype=np.dtype(dtype)))
    tf.count_nonzero = count_nonzero

    def argmax(x, axis=None, name=None, dimension=None, output_type=tf.int64):
        ax = axis if axis is not None else (dimension or 0)
        return Tensor(np.argmax(_t(x).a, axis=ax).astype(np.dtype(output_type)))
    tf.argmax = argmax

    def argmin(x, axis=None, name=None, dimension=None, output_type=tf.int64):
        ax = axis if axis is not None else (dimension or 0)
        return Tensor(np.argmin(_t(x).a, axis=ax).astype(np.dtype(output_type)))
    tf.argmin = argmin

    def where(condition, x=None, y=None, name=None):
        assert x is not None and y is not None
        return Tensor(np.where(_unwrap(condition), _unwrap(x), _unwrap(y)))
    tf.where = where

    def gather(params, indices, validate_indices=None, name=None, axis=0):
        return Tensor(np.take(_t(params).a, np.asarray(_unwrap(indices)), axis=axis))
    tf.gather = gather

    def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
        assert mode == "CONSTANT"
        p = np.asarray(_unwrap(paddings)).astype(np.int64)
        return Tensor(np.pad(_t(x).a, [tuple(r) for r in p.tolist()],
                             mode="constant", constant_values=constant_values))
    tf.pad = pad

    def cumsum(x, axis=0, exclusive=False, reverse=False, name=None):
        assert not exclusive and not reverse
        return Tensor(np.cumsum(_t(x).a, axis=axis, dtype=_t(x).a.dtype))
    tf.cumsum = cumsum
    tf.tuple = lambda tensors, name=None, control_inputs=None: [_t(t) for t in tensors]

    def add_n(inputs, name=None):
        acc = _t(inputs[0])
        for v in inputs[1:]:
            acc = acc + v                       # AddN accumulates its inputs in order
        return acc
    tf.add_n = add_n

    def reverse(x, axis, name=None):
        return Tensor(np.flip(_t(x).a, axis=tuple(int(a) for a in axis)))
    tf.reverse = reverse

    def scan(fn, elems, initializer=None, parallel_iterations=10, back_prop=True, swap_memory=False,
             infer_shape=True, reverse=False, name=None):
        a = _t(elems).a
        assert not reverse
        if a.shape[0] == 0:
            return Tensor(a.copy())
        out, acc, start = [], None, 0
        if initializer is None:
            acc, start = Tensor(a[0]), 1
            out.append(acc.a)
        else:
            acc = _t(initializer)
        for i in range(start, a.shape[0]):
            acc = _t(fn(acc, Tensor(a[i])))
            out.append(acc.a)
        return Tensor(np.stack(out).astype(a.dtype))
    tf.scan = scan

    def boolean_mask(x, mask, name=None, axis=None):
        return Tensor(_t(x).a[np.asarray(_unwrap(mask), dtype=bool)])
    tf.boolean_mask = boolean_mask

    def while_loop(cond, body, loop_vars, shape_invariants=None,
                   parallel_iterations=10, back_prop=True, swap_memory=False,
                   name=None, maximum_iterations=None):
        vs = list(loop_vars)
        while bool(_unwrap(cond(*vs))):
            vs = list(body(*vs))
        return vs
    tf.while_loop = while_loop

    def map_fn(fn, elems, dtype=None, parallel_iterations=10, back_prop=True,
               swap_memory=False, infer_shape=True, name=None):
        single = not isinstance(elems, (list, tuple))
        es = [_t(elems)] if single else [_t(e) for e in elems]
        n = es[0].a.shape[0]
        outs = []
        for i in range(n):
            arg = es[0][i] if single else type(elems)(e[i] for e in es)
            outs.append(fn(arg))
        if isinstance(outs[0], (list, tuple)):
            cols = list(zip(*outs))
            res = [Tensor(np.stack([_t(o).a for o in col])) for col in cols]
            return type(outs[0])(res) if isinstance(outs[0], tuple) else res
        return Tensor(np.stack([_t(o).a for o in outs]))
    tf.map_fn = map_fn

    class TensorArray:
        def __init__(self, dtype, size=None, dynamic_size=False, infer_shape=True, **kw):
            self._d = np.dtype(dtype)
            self._v = [None] * int(_unwrap(size))

        def write(self, i, v):
            self._v[int(_unwrap(i))] = np.asarray(_unwrap(v), dtype=self._d)
            return self

        def stack(self):
            return Tensor(np.stack(self._v) if self._v else np.zeros((0,), self._d))
    tf.TensorArray = TensorArray

    # ---- tf.nn ------------------------------------------------------------ #
    nn = types.ModuleType("tensorflow.nn")

    def top_k(x, k=1, sorted=True, name=None):
        """TopKV2: descending; equal values keep the lower index first."""
        a = _t(x).a
        k = int(_unwrap(k))
        assert k <= a.shape[-1], "top_k: k must be <= last dimension"
        with np.errstate(all="ignore"):
            order = np.argsort(-a, axis=-1, kind="stable")[..., :k]
        return (Tensor(np.take_along_axis(a, order, axis=-1)),
                Tensor(order.astype(np.int32)))
    nn.top_k = top_k

    def softmax(logits, axis=-1, name=None):
        a = _t(logits).a
        e = np.exp(a - np.max(a, axis=axis, keepdims=True))
        return Tensor((e / np.sum(e, axis=axis, keepdims=True)).astype(a.dtype))
    nn.softmax = softmax

    def moments(x, axes, shift=None, name=None, keep_dims=False):
        a = _t(x).a
        ax = tuple(int(v) for v in axes)
        mean = np.mean(a, axis=ax, keepdims=True, dtype=a.dtype)
        var = np.mean((a - mean) * (a - mean), axis=ax, keepdims=True, dtype=a.dtype)    # mean of squared difference
        if not keep_dims:
            mean, var = np.squeeze(mean, ax), np.squeeze(var, ax)
        return Tensor(mean), Tensor(var)
    nn.moments = moments

    def sparse_softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        x = _t(logits).a
        lab = np.asarray(_unwrap(labels)).astype(np.int64)
        z = x - np.max(x, axis=-1, keepdims=True)
        lse = np.log(np.sum(np.exp(z), axis=-1, dtype=x.dtype))
        return Tensor((lse - np.take_along_axis(z, lab[..., None], axis=-1)[..., 0]).astype(x.dtype))
    nn.sparse_softmax_cross_entropy_with_logits = sparse_softmax_cross_entropy_with_logits
    tf.nn = nn
    tf.pow = _bi(np.power)
    tf.div = lambda x, y, name=None: _t(x) / y
    # tf.contrib.slim.softmax is executed by det_clf_loss (utils/net_tools.py:571); the rest of contrib is mocked
    contrib = mock.MagicMock(name="tensorflow.contrib")
    contrib.slim.softmax = softmax
    tf.contrib = contrib

    # ---- tf.image --------------------------------------------------------- #
    image = types.ModuleType("tensorflow.image")

    def _nms_iou(b, i, j):
        f = np.float32
        ymin_i, ymax_i = min(b[i, 0], b[i, 2]), max(b[i, 0], b[i, 2])
        xmin_i, xmax_i = min(b[i, 1], b[i, 3]), max(b[i, 1], b[i, 3])
        ymin_j, ymax_j = min(b[j, 0], b[j, 2]), max(b[j, 0], b[j, 2])
        xmin_j, xmax_j = min(b[j, 1], b[j, 3]), max(b[j, 1], b[j, 3])
        area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
        area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
        if area_i <= 0 or area_j <= 0:
            return f(0.0)
        iymin, ixmin = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
        iymax, ixmax = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
        inter = f(max(f(iymax - iymin), f(0.0)) * max(f(ixmax - ixmin), f(0.0)))
        return f(inter / f(f(area_i + area_j) - inter))

    def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5,
                            score_threshold=float("-inf"), name=None):
        b = np.asarray(_unwrap(boxes), dtype=np.float32)
        s = np.asarray(_unwrap(scores), dtype=np.float32)
        max_out = int(_unwrap(max_output_size))
        thr = np.float32(_unwrap(iou_threshold))
        with np.errstate(all="ignore"):
            order = np.argsort(-s, kind="stable")
            selected = []
            for c in order:
                if len(selected) >= max_out:
                    break
                if not (s[c] > score_threshold):
                    break
                keep = True
                for j in reversed(selected):
                    if _nms_iou(b, c, j) > thr:
                        keep = False
                        break
                if keep:
                    selected.append(int(c))
        return Tensor(np.asarray(selected, dtype=np.int32))
    image.non_max_suppression = non_max_suppression
    tf.image = image

    # tf.train.Example & co., tf.python_io.TFRecordWriter, tf.gfile: the dataset converter (dataset/pascalvoc_to_tfrecords.py)
    from . import example_proto
    example_proto.attach(tf)

    # anything else (tf.contrib, tf.app, tf.summary, tf.GraphKeys, ...) is
    # imported-but-not-executed on the hot path: hand out mocks.
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name="tensorflow." + name)
        setattr(tf, name, m)
        return m
    tf.__getattr__ = __getattr__
    return tf


# --------------------------------------------------------------------------- #
# Import machinery
# --------------------------------------------------------------------------- #
class _MockLoader(importlib.abc.Loader):
    def __init__(self, real):
        self._real = real

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        m.__name__ = spec.name
        m.__loader__ = self
        # `from <mock pkg> import child` looks the attribute up first: make real
        # children win over auto-created mock attributes.
        for full, mod in self._real.items():
            parent, _, child = full.rpartition(".")
            if parent == spec.name:
                setattr(m, child, mod)
        return m

    def exec_module(self, module):
        pass


class _RealLoader(importlib.abc.Loader):
    def __init__(self, module):
        self._m = module

    def create_module(self, spec):
        self._m.__spec__ = spec
        self._m.__loader__ = self
        return self._m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def __init__(self, real):
        self._real = real

    def find_spec(self, name, path, target=None):
        if name in self._real:
            return importlib.machinery.ModuleSpec(
                name, _RealLoader(self._real[name]), is_package=True)
        if name == "tensorflow" or name.startswith("tensorflow."):
            return importlib.machinery.ModuleSpec(name, _MockLoader(self._real), is_package=True)
        return None


_installed = None


def install():
    """Put the shim on sys.meta_path (idempotent); returns the `tensorflow` module."""
    global _installed
    if _installed is not None:
        return _installed
    if "tensorflow" in sys.modules:
        raise RuntimeError("a real tensorflow is already imported; shim not needed")
    tf = _build_tf_module()
    # tensorflow.python.ops.math_ops is *executed* by tf_extended/math.py:safe_divide
    math_ops = types.ModuleType("tensorflow.python.ops.math_ops")
    math_ops.__path__ = []
    math_ops.greater = tf.greater
    math_ops.divide = tf.divide
    math_ops.mul = tf.multiply
    math_ops.to_float = lambda x, name=None: tf.cast(x, np.float32)
    math_ops.to_int64 = lambda x, name=None: tf.cast(x, np.int64)
    # tensorflow.python.framework.ops is *executed* by tf_extended/math.py:cummax and metrics.py
    import contextlib
    fw_ops = types.ModuleType("tensorflow.python.framework.ops")
    fw_ops.__path__ = []

    @contextlib.contextmanager
    def _ops_name_scope(name=None, default_name=None, values=None):
        yield name or default_name
    fw_ops.name_scope = _ops_name_scope
    fw_ops.convert_to_tensor = lambda v, dtype=None, name=None: tf.convert_to_tensor(v, dtype=dtype)
    fw_ops.GraphKeys = mock.MagicMock(name="GraphKeys")
    fw_ops.add_to_collections = lambda *a, **k: None
    real = {
        "tensorflow.python.framework.ops": fw_ops,
        "tensorflow": tf,
        "tensorflow.nn": tf.nn,
        "tensorflow.image": tf.image,
        "tensorflow.python.ops.math_ops": math_ops,
    }
    sys.meta_path.insert(0, _Finder(real))
    _installed = tf
    return tf
