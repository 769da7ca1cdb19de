"""Reverse mode through Op.L_op (pytensor.gradient)."""
from .graph.basic import Variable, ancestors_nodes


class DisconnectedType:
    ndim = 0
    dtype = None

    def __call__(self, name=None):
        return Variable(self, name=name)

    def __repr__(self):
        return "DisconnectedType"


class NullType:
    ndim = 0
    dtype = None

    def __init__(self, why_null="(no reason given)"):
        self.why_null = why_null

    def __call__(self, name=None):
        return Variable(self, name=name)

    def __repr__(self):
        return f"NullType({self.why_null})"


disconnected_type = DisconnectedType()


def grad_undefined(op, x_pos, x, comment=""):
    return NullType(f"gradient of {type(op).__name__} w.r.t. input {x_pos} is undefined. {comment}")()


class DisconnectedInputError(ValueError):
    pass


class NullTypeGradError(TypeError):
    pass


def grad(cost, wrt, disconnected_inputs="raise", null_gradients="raise"):
    """d cost / d wrt for a scalar ``cost``."""
    from . import tensor as pt

    single = not isinstance(wrt, (list, tuple))
    wrts = [wrt] if single else list(wrt)
    if cost.type.ndim != 0:
        raise TypeError("cost must be a scalar")
    gmap = {id(cost): pt.ones_like(cost)}
    for node in reversed(ancestors_nodes([cost])):
        ogs = [gmap.get(id(o)) for o in node.outputs]
        if all(g is None for g in ogs):
            continue
        ogs = [disconnected_type() if g is None else g for g in ogs]
        igs = node.op.L_op(node.inputs, node.outputs, ogs)
        if len(igs) != len(node.inputs):
            raise ValueError(f"{type(node.op).__name__}.L_op returned {len(igs)} gradients for {len(node.inputs)} inputs")
        pattern = node.op.connection_pattern(node)
        out_connected = [not isinstance(g.type, DisconnectedType) for g in ogs]
        for i, (inp, ig) in enumerate(zip(node.inputs, igs)):
            connected = any(c and oc for c, oc in zip(pattern[i], out_connected))
            if not connected and not isinstance(ig.type, DisconnectedType):
                # the check pytensor.gradient makes: a disconnected input must get a DisconnectedType cotangent
                raise TypeError(
                    f"{type(node.op).__name__}.L_op returned {ig.type} for input {i}, expected DisconnectedType instance "
                    "based on the output of the op's connection_pattern method")
            if isinstance(ig.type, DisconnectedType):
                continue
            if isinstance(ig.type, NullType):
                gmap[id(inp)] = ig
                continue
            prev = gmap.get(id(inp))
            gmap[id(inp)] = ig if prev is None else pt.add(prev, ig)
    out = []
    for w in wrts:
        g = gmap.get(id(w))
        if g is None:
            if disconnected_inputs == "raise":
                raise DisconnectedInputError(f"{w} is not part of the graph of the cost")
            g = pt.zeros_like(w)
        elif isinstance(g.type, NullType):
            if null_gradients == "raise":
                raise NullTypeGradError(g.type.why_null)
        out.append(g)
    return out[0] if single else out
