"""Variable / Constant / Apply (pytensor.graph.basic)."""


class Variable:
    def __init__(self, type, owner=None, index=None, name=None):
        self.type, self.owner, self.index, self.name = type, owner, index, name

    # the tensor conveniences PyTensor puts on TensorVariable
    @property
    def ndim(self):
        return self.type.ndim

    @property
    def dtype(self):
        return self.type.dtype

    @property
    def shape(self):
        from ..tensor import shape_of

        return shape_of(self)

    def sum(self, axis=None):
        from ..tensor import sum as _sum

        return _sum(self, axis=axis)

    def reshape(self, shape):
        from ..tensor import reshape

        return reshape(self, shape)

    def __mul__(self, other):
        from ..tensor import mul

        return mul(self, other)

    __rmul__ = __mul__

    def __add__(self, other):
        from ..tensor import add

        return add(self, other)

    __radd__ = __add__

    def __repr__(self):
        return f"<{type(self).__name__} {self.name or ''} {self.type}>"


class Constant(Variable):
    def __init__(self, type, data, name=None):
        super().__init__(type, None, None, name)
        self.data = data


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, o in enumerate(self.outputs):
            o.owner, o.index = self, i

    @property
    def nin(self):
        return len(self.inputs)


def ancestors_nodes(outputs):
    """Apply nodes the outputs depend on, in topological order."""
    order, seen = [], set()

    def visit(v):
        node = v.owner
        if node is None or id(node) in seen:
            return
        seen.add(id(node))
        for i in node.inputs:
            visit(i)
        order.append(node)

    for o in outputs:
        visit(o)
    return order
