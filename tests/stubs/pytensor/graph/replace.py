"""vectorize_graph and the per-Op dispatch _vectorize_node (pytensor.graph.replace)."""
from functools import singledispatch

from .basic import Constant, ancestors_nodes


@singledispatch
def _vectorize_node(op, node, *batched_inputs):
    # PyTensor's default wraps the Op in Blockwise (a Python loop over the batch); the stub refuses instead, so that a test
    # passes only when the Op registered a natively batched replacement
    raise NotImplementedError(f"no _vectorize_node registered for {type(op).__name__}")


def vectorize_node(node, *batched_inputs):
    return _vectorize_node(node.op, node, *batched_inputs)


def vectorize_graph(outputs, replace):
    """Rebuild ``outputs`` with the variables in ``replace`` (old -> new with extra leading dims) substituted."""
    single = not isinstance(outputs, (list, tuple))
    outs = [outputs] if single else list(outputs)
    memo = {id(k): v for k, v in replace.items()}

    def new_of(v):
        return memo.get(id(v), v)

    for node in ancestors_nodes(outs):
        new_inputs = [new_of(i) for i in node.inputs]
        if all(a is b for a, b in zip(new_inputs, node.inputs)):
            continue
        new_node = vectorize_node(node, *new_inputs)
        for o, no in zip(node.outputs, new_node.outputs):
            memo[id(o)] = no
    res = [new_of(o) for o in outs]
    return res[0] if single else res


__all__ = ["_vectorize_node", "vectorize_node", "vectorize_graph", "Constant"]
