"""The few tensor functions the Op layer and its tests use (pytensor.tensor)."""
import numpy as np

from ..graph.basic import Apply, Constant, Variable
from ..graph.op import Op


class TensorType:
    def __init__(self, dtype, shape):
        self.dtype = str(np.dtype(dtype))
        self.shape = tuple(None if s is None else int(s) for s in shape)

    @property
    def ndim(self):
        return len(self.shape)

    def __call__(self, name=None):
        return Variable(TensorType(self.dtype, self.shape), name=name)

    def __eq__(self, other):
        return isinstance(other, TensorType) and (self.dtype, self.shape) == (other.dtype, other.shape)

    def __hash__(self):
        return hash((self.dtype, self.shape))

    def __repr__(self):
        return f"Tensor({self.dtype}, {self.shape})"


def tensor(name=None, *, dtype=None, shape=None):
    from .. import config

    return Variable(TensorType(dtype or config.floatX, shape), name=name)


def matrix(name=None, dtype=None, shape=(None, None)):
    return tensor(name, dtype=dtype, shape=shape)


def vector(name=None, dtype=None, shape=(None,)):
    return tensor(name, dtype=dtype, shape=shape)


def tensor3(name=None, dtype=None, shape=(None, None, None)):
    return tensor(name, dtype=dtype, shape=shape)


def as_tensor_variable(x, name=None, dtype=None):
    if isinstance(x, Variable):
        return x
    a = np.asarray(x) if dtype is None else np.asarray(x, dtype=dtype)
    return Constant(TensorType(a.dtype, a.shape), a, name=name)


as_tensor = as_tensor_variable
constant = as_tensor_variable


class _Elementwise(Op):
    """out = fn(*inputs) with NumPy broadcasting; L_op written per subclass."""

    fn = None

    def make_node(self, *inputs):
        ins = [as_tensor_variable(x) for x in inputs]
        nd = max(i.type.ndim for i in ins)
        shape = []
        for k in range(nd):
            dims = [i.type.shape[k - nd + i.type.ndim] for i in ins if k - nd + i.type.ndim >= 0]
            known = [d for d in dims if d not in (None, 1)]
            shape.append(known[0] if known else (1 if dims and all(d == 1 for d in dims) else None))
        dtype = str(np.result_type(*[i.type.dtype for i in ins]))
        return Apply(self, ins, [Variable(TensorType(dtype, shape))])

    def perform(self, node, inputs, output_storage, params=None):
        output_storage[0][0] = np.asarray(type(self).fn(*inputs))


def _unbroadcast(g, like):
    """Sum a cotangent over the axes ``like`` was broadcast along."""
    extra = g.type.ndim - like.type.ndim
    if extra > 0:
        g = sum(g, axis=tuple(range(extra)))
    return g


class Add(_Elementwise):
    fn = staticmethod(lambda a, b: a + b)

    def L_op(self, inputs, outputs, ogs):
        return [_unbroadcast(ogs[0], inputs[0]), _unbroadcast(ogs[0], inputs[1])]


class Mul(_Elementwise):
    fn = staticmethod(lambda a, b: a * b)

    def L_op(self, inputs, outputs, ogs):
        return [_unbroadcast(mul(ogs[0], inputs[1]), inputs[0]), _unbroadcast(mul(ogs[0], inputs[0]), inputs[1])]


def add(a, b):
    return Add()(a, b)


def mul(a, b):
    return Mul()(a, b)


class Sum(Op):
    __props__ = ("axis",)

    def __init__(self, axis=None):
        self.axis = axis if axis is None or isinstance(axis, tuple) else (axis,)

    def make_node(self, x):
        x = as_tensor_variable(x)
        axes = tuple(range(x.type.ndim)) if self.axis is None else tuple(a % x.type.ndim for a in self.axis)
        shape = [s for k, s in enumerate(x.type.shape) if k not in axes]
        return Apply(self, [x], [Variable(TensorType(x.type.dtype, shape))])

    def perform(self, node, inputs, output_storage, params=None):
        output_storage[0][0] = np.asarray(inputs[0].sum(axis=self.axis))

    def L_op(self, inputs, outputs, ogs):
        return [FillLike(self.axis)(inputs[0], ogs[0])]


class FillLike(Op):
    """Broadcast the cotangent of a Sum back to the shape of the summed input."""

    __props__ = ("axis",)

    def __init__(self, axis):
        self.axis = axis

    def make_node(self, like, g):
        return Apply(self, [like, as_tensor_variable(g)], [like.type()])

    def perform(self, node, inputs, output_storage, params=None):
        like, g = inputs
        g = np.asarray(g)
        if self.axis is not None:
            g = np.expand_dims(g, tuple(a % like.ndim for a in self.axis))
        output_storage[0][0] = np.broadcast_to(g, like.shape).astype(like.dtype).copy()


def sum(x, axis=None):  # noqa: A001
    return Sum(axis)(x)


class _Like(Op):
    __props__ = ("value",)

    def __init__(self, value):
        self.value = value

    def make_node(self, x):
        x = as_tensor_variable(x)
        return Apply(self, [x], [x.type()])

    def perform(self, node, inputs, output_storage, params=None):
        output_storage[0][0] = np.full_like(inputs[0], self.value)

    def L_op(self, inputs, outputs, ogs):
        from ..gradient import disconnected_type

        return [disconnected_type()]

    def connection_pattern(self, node):
        return [[False]]


def zeros_like(x):
    return _Like(0.0)(x)


def ones_like(x):
    return _Like(1.0)(x)


def zeros(shape, dtype=None):
    from .. import config

    return as_tensor_variable(np.zeros(shape, dtype=dtype or config.floatX))


def eye(n, dtype=None):
    from .. import config

    return as_tensor_variable(np.eye(n, dtype=dtype or config.floatX))


class _ShapeTuple(tuple):
    """``x.shape`` of a symbolic variable: static entries as ints, unknown ones as ``ShapeEntry`` markers that
    broadcast_to / reshape resolve at run time."""


class ShapeEntry:
    def __init__(self, var, axis):
        self.var, self.axis = var, axis


def shape_of(x):
    return _ShapeTuple(s if s is not None else ShapeEntry(x, k) for k, s in enumerate(x.type.shape))


class BroadcastTo(Op):
    """broadcast_to(x, shape) where the entries of shape are ints or ShapeEntry markers (their variables ride along as
    inputs, so that the evaluator has them)."""

    __props__ = ("spec",)

    def __init__(self, spec):
        self.spec = spec  # tuple of ints or ("dyn", input position, axis)

    def make_node(self, x, *refs):
        x = as_tensor_variable(x)
        shape = [s if isinstance(s, int) else None for s in self.spec]
        return Apply(self, [x, *refs], [Variable(TensorType(x.type.dtype, shape))])

    def perform(self, node, inputs, output_storage, params=None):
        shape = tuple(s if isinstance(s, int) else inputs[s[1]].shape[s[2]] for s in self.spec)
        output_storage[0][0] = np.broadcast_to(inputs[0], shape).copy()


def broadcast_to(x, shape):
    refs, spec = [], []
    for s in shape:
        if isinstance(s, ShapeEntry):
            refs.append(s.var)
            spec.append(("dyn", len(refs), s.axis))
        else:
            spec.append(int(s))
    return BroadcastTo(tuple(spec))(x, *refs)


class Reshape(Op):
    __props__ = ("shape",)

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)

    def make_node(self, x):
        x = as_tensor_variable(x)
        return Apply(self, [x], [Variable(TensorType(x.type.dtype, [None if s == -1 else s for s in self.shape]))])

    def perform(self, node, inputs, output_storage, params=None):
        output_storage[0][0] = np.reshape(inputs[0], self.shape)


def reshape(x, shape):
    return Reshape(shape)(x)
