"""Stationary initialisation on the GPU -- host-side mirror of what the reference's models
call right before the filter:

* ``models/SARIMAX.py:369-381``, ``models/VARMAX.py:379-393``:
  ``x0 = pt.linalg.solve(I - T, c)``, ``P0 = solve_discrete_lyapunov(T, matrix_dot(R, Q, R.T), method=...)``
* ``models/ETS.py:453-466``: ``P0 = solve_discrete_lyapunov(T_stationary, matrix_dot(R, Q, R.T))``

``solve_discrete_lyapunov(A, Q)`` has the signature of ``pytensor.tensor.slinalg.solve_discrete_lyapunov`` (the ``method``
argument is accepted and ignored: there is one GPU method, doubling, see ``csrc/bkf_stationary.cu``); a maintainer swaps
the import in the three model files.  Eager inputs (NumPy / torch CUDA, optionally with a leading batch axis) run
immediately; PyTensor variables build a :class:`~pymc_experimental_b200.pytensor_ops.StationaryInitOp` node.
No CPU fallback.
"""

from __future__ import annotations

import numpy as np

from .engine import STATUS_NOT_STATIONARY, default_engine


def _is_symbolic(*xs):
    try:
        from pytensor.graph.basic import Variable
    except ImportError:
        return False
    return any(isinstance(x, Variable) for x in xs)


def stationary_initialization(c, T, R, Q, device: int = 0):
    """``(x0, P0)`` of ``_stationary_initialization`` (models/SARIMAX.py:369-381).  Raises if T is not stationary."""
    if _is_symbolic(c, T, R, Q):
        from . import pytensor_ops

        return pytensor_ops.StationaryInitOp(device)(c, T, R, Q)
    a0, P0, status = default_engine(device).stationary_init(c, T, R, Q)
    if np.any(np.asarray(status.cpu() if hasattr(status, "cpu") else status) & STATUS_NOT_STATIONARY):
        raise np.linalg.LinAlgError("stationary initialisation: the transition matrix is not stationary "
                                    "(spectral radius >= 1)")
    return a0, P0


def solve_discrete_lyapunov(A, Q, method: str | None = None, device: int = 0):
    """Solution X of ``A X A^T - X + Q = 0`` (``pytensor.tensor.slinalg.solve_discrete_lyapunov``)."""
    m = int(A.shape[-1])
    if _is_symbolic(A, Q):
        import pytensor.tensor as pt

        from . import pytensor_ops

        return pytensor_ops.StationaryInitOp(device)(pt.zeros((m,), dtype=A.dtype), A, pt.eye(m, dtype=A.dtype), Q)[1]
    xp_zeros = np.zeros(m)
    eye = np.eye(m)
    if type(A).__module__.startswith("torch"):
        import torch

        xp_zeros = torch.zeros(m, dtype=A.dtype, device=A.device)
        eye = torch.eye(m, dtype=A.dtype, device=A.device)
    return stationary_initialization(xp_zeros, A, eye, Q, device)[1]
