"""A protocol stub of PyTensor -- TEST INFRASTRUCTURE ONLY (never imported by the package, never shipped).

PyTensor cannot be installed in this project's containers (no wheel, no network), so the Op layer of
``pymc_experimental_b200.pytensor_ops`` is exercised against this stand-in.  It implements, from the public PyTensor API
as the SURVEY's Appendix C records it, exactly the parts an ``Op`` author touches:

* graph objects: ``Variable`` / ``Constant`` / ``Apply``, ``Op.__call__`` -> ``make_node``, ``__props__`` equality;
* ``pytensor.function(inputs, outputs)``: a Python evaluator that walks the graph and calls ``Op.perform(node, inputs,
  output_storage)`` once per reachable node (unreachable nodes are pruned, as PyTensor does);
* ``pytensor.grad`` / ``pytensor.gradient``: reverse mode through ``Op.L_op(inputs, outputs, output_grads)`` with
  ``DisconnectedType`` / ``NullType`` cotangents and the consistency check between ``connection_pattern`` and what
  ``L_op`` returns for disconnected inputs (the check ``pytensor.gradient`` makes);
* ``pytensor.graph.replace.vectorize_graph`` with the ``_vectorize_node`` single dispatch.

Only the handful of tensor functions the Op layer and its tests use exist (``pytensor.tensor``).  Anything else raises.
"""
from types import SimpleNamespace

config = SimpleNamespace(floatX="float64")

from . import gradient, graph, tensor  # noqa: E402
from .compile import function  # noqa: E402
from .gradient import grad  # noqa: E402

__all__ = ["config", "function", "grad", "gradient", "graph", "tensor"]
__version__ = "0.0-stub"
