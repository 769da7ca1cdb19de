"""pytensor.function: evaluate a graph by calling Op.perform on every node the outputs need."""
import numpy as np

from .graph.basic import Constant, ancestors_nodes


class Function:
    def __init__(self, inputs, outputs):
        self.single = not isinstance(outputs, (list, tuple))
        self.inputs = list(inputs)
        self.outputs = [outputs] if self.single else list(outputs)
        self.nodes = ancestors_nodes(self.outputs)
        self.performed = []  # Op instances run by the last call, in order

    def __call__(self, *values):
        if len(values) != len(self.inputs):
            raise TypeError(f"expected {len(self.inputs)} inputs, got {len(values)}")
        env = {}
        for v, x in zip(self.inputs, values):
            env[id(v)] = np.asarray(x, dtype=v.type.dtype)
        self.performed = []

        def value(v):
            if isinstance(v, Constant):
                return v.data
            if id(v) not in env:
                raise ValueError(f"missing input {v}")
            return env[id(v)]

        for node in self.nodes:
            ins = [value(i) for i in node.inputs]
            storage = [[None] for _ in node.outputs]
            node.op.perform(node, ins, storage)
            self.performed.append(node.op)
            for o, st in zip(node.outputs, storage):
                if st[0] is None:
                    raise RuntimeError(f"{type(node.op).__name__}.perform left an output empty")
                env[id(o)] = st[0]
        res = [value(o) for o in self.outputs]
        return res[0] if self.single else res


def function(inputs, outputs, **kwargs):
    return Function(inputs, outputs)
