"""Op base class (pytensor.graph.op.Op)."""


class Op:
    __props__ = ()

    def _props(self):
        return tuple(getattr(self, k) for k in self.__props__)

    def __eq__(self, other):
        return type(self) is type(other) and self._props() == other._props()

    def __hash__(self):
        return hash((type(self), self._props()))

    def __call__(self, *inputs, **kwargs):
        node = self.make_node(*inputs, **kwargs)
        return node.outputs[0] if len(node.outputs) == 1 else list(node.outputs)

    def make_node(self, *inputs):
        raise NotImplementedError

    def perform(self, node, inputs, output_storage, params=None):
        raise NotImplementedError

    def L_op(self, inputs, outputs, output_grads):
        raise NotImplementedError(f"{type(self).__name__} has no L_op")

    def connection_pattern(self, node):
        return [[True] * len(node.outputs) for _ in node.inputs]

    def infer_shape(self, fgraph, node, input_shapes):
        raise NotImplementedError
