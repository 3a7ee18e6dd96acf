"""A torch-backed stand-in for the handful of `tensorflow.compat.v1` calls the reference makes
(TEST INFRASTRUCTURE; used only by oracle/gen_golden.py in a container where /root/reference is
mounted).

TensorFlow cannot be installed here, so the reference's UNMODIFIED model files (model_1..4.py,
model.py) are executed over this shim to produce the golden vectors under tests/golden/.  The
shim is a lazy graph: every `tf.*` call returns a Node; `Session.run(fetches, feed_dict)`
evaluates the requested nodes with torch (float64 by default, so golden vectors carry no fp32
noise) and autograd provides the gradients for `AdamOptimizer.minimize`.

Semantics implemented from TF1's documented behaviour:
  tf.matmul          batched matrix product            tf.transpose(x, perm)   permute
  tf.nn.l2_loss      sum(x^2)/2                        tf.nn.softmax(dim=)     softmax over `dim`
  tf.nn.softmax_cross_entropy_with_logits(dim=d)       -sum_d labels*log_softmax(logits)
  tf.truncated_normal  normal, re-drawn beyond 2 sigma tf.train.AdamOptimizer  lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
                                                       p -= lr_t*m/(sqrt(v)+eps), m/v un-corrected
Variable names follow TF's `scope/name` with `_k` suffixes on reuse, so
`tf.global_variables()` order == creation order (what the flat parameter blob relies on).
"""
from __future__ import annotations

import contextlib
import math
import sys
import types

import numpy as np
import torch

DTYPE = torch.float64
float32 = "float32"
_state = types.SimpleNamespace(variables=[], scopes=[], names={}, gen=torch.Generator().manual_seed(0))


def set_seed(seed: int):
    _state.gen = torch.Generator().manual_seed(seed)


class Node:
    __array_ufunc__ = None          # `ndarray - Node` defers to Node.__rsub__ (as tf.Tensor does), model.py:372

    def __init__(self, fn, inputs=(), name=None):
        self.fn, self.inputs, self.name = fn, tuple(inputs), name
        with torch.no_grad():
            metas = [i.meta if isinstance(i, Node) else i for i in self.inputs]
            self.meta = fn(*metas) if fn is not None else None

    @property
    def shape(self):
        return _Shape(self.meta.shape)

    def get_shape(self):
        return self.shape

    def _bin(self, other, op, rev=False):
        o = other if isinstance(other, (Node, int, float)) else constant(other)
        a, b = (o, self) if rev else (self, o)
        return Node(op, (a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Node(torch.neg, (self,))


class _Dim(int):
    @property
    def value(self):
        return int(self)


class _Shape(tuple):
    def __new__(cls, s):
        return super().__new__(cls, [_Dim(d) for d in s])

    def as_list(self):
        return [int(d) for d in self]


class _Const(Node):
    def __init__(self, t, name=None):
        self.fn, self.inputs, self.name, self.t = None, (), name, t
        self.meta = torch.empty(t.shape, dtype=DTYPE, device="meta")


def constant(v, dtype=None, name=None):
    return _Const(torch.as_tensor(v, dtype=DTYPE), name)


class Placeholder(Node):
    def __init__(self, shape, name):
        self.fn, self.inputs, self.name = None, (), name
        self.meta = torch.empty(tuple(shape), dtype=DTYPE, device="meta")


def placeholder(dtype, shape=None, name=None):
    return Placeholder(shape, name)


class Variable(Node):
    def __init__(self, initial_value, name=None, dtype=None, trainable=True):
        init = initial_value() if callable(initial_value) else initial_value
        if isinstance(init, Node):
            init = _eval(init, {}, {})
        self.value = torch.as_tensor(init, dtype=DTYPE).clone().requires_grad_(True)
        self.initial = self.value.detach().clone()
        scope = "/".join(_state.scopes)
        base = (scope + "/" if scope else "") + (name or "Variable")
        k = _state.names.get(base, 0)
        _state.names[base] = k + 1
        self.name = (base if k == 0 else f"{base}_{k}") + ":0"
        self.fn, self.inputs = None, ()
        self.meta = torch.empty(self.value.shape, dtype=DTYPE, device="meta")
        self.trainable = trainable
        _state.variables.append(self)


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    shape = [int(s) for s in shape]
    n = int(np.prod(shape))
    v = torch.randn(n, generator=_state.gen, dtype=torch.float64)
    bad = v.abs() > 2
    while bad.any():
        v[bad] = torch.randn(int(bad.sum()), generator=_state.gen, dtype=torch.float64)
        bad = v.abs() > 2
    return (mean + stddev * v).reshape(shape)


def zeros(shape, dtype=None, name=None):
    return torch.zeros([int(s) for s in shape], dtype=DTYPE)


def ones(shape, dtype=None, name=None):
    return torch.ones([int(s) for s in shape], dtype=DTYPE)


def _n(x):
    return x if isinstance(x, Node) else constant(x)


def matmul(a, b, name=None): return Node(torch.matmul, (_n(a), _n(b)))
def multiply(a, b, name=None): return Node(torch.mul, (_n(a), _n(b)))
def divide(a, b, name=None): return Node(torch.div, (_n(a), _n(b)))
def transpose(a, perm=None, name=None): return Node(lambda t: t.permute(*perm) if perm is not None else t.T, (_n(a),))
def reshape(a, shape, name=None):
    shp = [int(s) for s in shape]
    return Node(lambda t: t.reshape(shp), (_n(a),))
def concat(values, axis, name=None):
    vals = [_n(v) for v in values]
    return Node(lambda *ts: torch.cat(ts, axis), vals)
def tile(a, multiples, name=None): return Node(lambda t: t.repeat(*[int(m) for m in multiples]), (_n(a),))
def sqrt(a, name=None): return Node(torch.sqrt, (_n(a),))
def square(a, name=None): return Node(torch.square, (_n(a),))
def pow(a, b, name=None): return Node(lambda t, e: torch.pow(t, e.to(DTYPE) if isinstance(e, torch.Tensor) else e), (_n(a), _n(b)))  # noqa: A001
def cast(a, dtype, name=None): return _n(a)
def argmax(a, axis=None, name=None): return Node(lambda t: torch.argmax(t, axis), (_n(a),))
def matrix_diag(a, name=None): return Node(torch.diag_embed, (_n(a),))


def _reduce(fn):
    def f(a, axis=None, keepdims=False, name=None, reduction_indices=None):
        ax = axis if axis is not None else reduction_indices
        if ax is None:
            return Node(lambda t: fn(t), (_n(a),))
        ax_ = tuple(ax) if isinstance(ax, (list, tuple)) else ax
        if fn in (torch.amax, torch.amin):
            return Node(lambda t: fn(t, dim=ax_, keepdim=keepdims), (_n(a),))
        return Node(lambda t: fn(t, dim=ax_, keepdim=keepdims), (_n(a),))
    return f


reduce_mean = _reduce(torch.mean)
reduce_sum = _reduce(torch.sum)
reduce_max = _reduce(torch.amax)
reduce_min = _reduce(torch.amin)


class _NN:
    @staticmethod
    def relu(a, name=None): return Node(torch.relu, (_n(a),))

    @staticmethod
    def softmax(a, axis=None, name=None, dim=None):
        d = dim if dim is not None else (axis if axis is not None else -1)
        return Node(lambda t: torch.softmax(t, d), (_n(a),))

    @staticmethod
    def l2_loss(a, name=None): return Node(lambda t: 0.5 * (t * t).sum(), (_n(a),))

    @staticmethod
    def softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, dim=-1, name=None):
        return Node(lambda lg, lb: -(lb * torch.log_softmax(lg, dim)).sum(dim), (_n(logits), _n(labels)))


nn = _NN()


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _state.scopes.append(name)
    try:
        yield name
    finally:
        _state.scopes.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


def global_variables(): return list(_state.variables)
def trainable_variables(): return [v for v in _state.variables if v.trainable]


def reset_default_graph():
    _state.variables.clear(); _state.scopes.clear(); _state.names.clear()


def disable_v2_behavior(): pass


class _Op:
    """A stateful op (initializer / train step)."""
    def __init__(self, run): self.run = run


def global_variables_initializer():
    def run(sess, feed):
        for v in _state.variables:
            v.value = v.initial.clone().requires_grad_(True)
    return _Op(run)


class _Summary:
    @staticmethod
    def scalar(*a, **k): return None
    @staticmethod
    def histogram(*a, **k): return None
    @staticmethod
    def merge_all(): return _Op(lambda sess, feed: b"")
    class FileWriter:
        def __init__(self, *a, **k): pass
        def add_summary(self, *a, **k): pass
        def close(self): pass


summary = _Summary()


class _Adam:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.slots = {}
        self.last_grads = None

    def minimize(self, loss, var_list=None):
        def run(sess, feed):
            vs = var_list or trainable_variables()
            val = _eval(loss, feed, {})
            grads = torch.autograd.grad(val, [v.value for v in vs], allow_unused=True)
            self.t += 1
            lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
            self.last_grads = [g.detach().clone() if g is not None else torch.zeros_like(v.value) for g, v in zip(grads, vs)]
            with torch.no_grad():
                for v, g in zip(vs, self.last_grads):
                    m, s = self.slots.get(id(v), (torch.zeros_like(g), torch.zeros_like(g)))
                    m = self.b1 * m + (1 - self.b1) * g
                    s = self.b2 * s + (1 - self.b2) * g * g
                    self.slots[id(v)] = (m, s)
                    v.value -= lr_t * m / (s.sqrt() + self.eps)
        op = _Op(run)
        op.optimizer = self
        return op


class _Saver:
    def __init__(self, *a, **k): self.store = None
    def save(self, sess, path, global_step=None): self.store = [v.value.detach().clone() for v in _state.variables]
    def restore(self, sess, path):
        for v, s in zip(_state.variables, self.store or []):
            v.value = s.clone().requires_grad_(True)


class _Train:
    AdamOptimizer = _Adam
    Saver = _Saver
    @staticmethod
    def get_checkpoint_state(d): return None


train = _Train()


def _eval(node, feed, memo):
    if not isinstance(node, Node):
        return node
    if id(node) in memo:
        return memo[id(node)]
    if isinstance(node, _Const):
        out = node.t
    elif isinstance(node, Variable):
        out = node.value
    elif isinstance(node, Placeholder):
        if node not in feed:
            raise KeyError(f"placeholder {node.name} not fed")
        out = torch.as_tensor(np.asarray(feed[node]), dtype=DTYPE)
        assert tuple(out.shape) == tuple(node.meta.shape), (node.name, out.shape, node.meta.shape)
    else:
        out = node.fn(*[_eval(i, feed, memo) for i in node.inputs])
    memo[id(node)] = out
    return out


class Session:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def close(self): pass

    def run(self, fetches, feed_dict=None):
        feed = feed_dict or {}
        memo = {}
        ops = []

        def one(f):
            if isinstance(f, _Op):
                ops.append(f)
                return None
            if isinstance(f, Node):
                return _eval(f, feed, memo).detach().numpy().copy()
            return f

        if isinstance(f := fetches, (list, tuple)):
            res = [one(x) for x in f]
        else:
            res = one(f)
        for op in ops:                     # values are fetched BEFORE the update, as in TF1
            op.run(self, feed)
        return res


class _Flags:
    FLAGS = types.SimpleNamespace()


class _App:
    flags = _Flags()
    @staticmethod
    def run(main=None, argv=None):
        m = main or sys.modules["__main__"].main
        m(sys.argv)


app = _App()


def install():
    """Make `import tensorflow.compat.v1 as tf` resolve to this module."""
    from importlib.machinery import ModuleSpec
    me = sys.modules[__name__]
    pkg = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    pkg.__spec__ = ModuleSpec("tensorflow", None, is_package=True)       # torch._dynamo probes find_spec("tensorflow")
    compat.__spec__ = ModuleSpec("tensorflow.compat", None, is_package=True)
    pkg.__path__, compat.__path__ = [], []
    pkg.compat = compat
    compat.v1 = me
    for k, v in vars(me).items():          # the legacy model.py does `import tensorflow as tf` (model.py:6)
        if not k.startswith("_") and k not in ("sys", "types", "np", "torch", "math"):
            setattr(pkg, k, v)
    sys.modules["tensorflow"] = pkg
    sys.modules["tensorflow.compat"] = compat
    sys.modules["tensorflow.compat.v1"] = me
