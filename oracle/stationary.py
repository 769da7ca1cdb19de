"""CPU oracle for the stationary initialisation.  TEST INFRASTRUCTURE ONLY.

Reference call sites (paths relative to pymc-devs/pymc-experimental):
  models/SARIMAX.py:369-381   x0 = pt.linalg.solve(I - T, c, assume_a="gen");  P0 = solve_discrete_lyapunov(T, R Q R^T)
  models/VARMAX.py:379-393    the same, method = "direct" if k_states < 10 else "bilinear"
  models/ETS.py:453-466       P0 = solve_discrete_lyapunov(T_stationary, R Q R^T)

``solve_discrete_lyapunov`` lives in PyTensor (``pytensor.tensor.slinalg``; third party, not installable here -- parity
unpinned against the reference's own output).  Its two methods are restated from its published definition:
"direct" solves the Kronecker system (I - A (x) A) vec(X) = vec(Q); "bilinear" calls
``scipy.linalg.solve_discrete_lyapunov``.  Both are checked against each other below; the torch version of "direct"
provides the gradient oracle (its autograd result equals the L_op of PyTensor's Op: Xbar = lyap(A^T, dX),
Abar = Xbar A X^T + Xbar^T A X, Qbar = Xbar).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg
import torch


def stationary_init_np(c, T, R, Q, method="direct"):
    """(a0, P0) for one model; numpy / scipy."""
    m = T.shape[0]
    a0 = np.linalg.solve(np.eye(m) - T, c)
    RQR = R @ Q @ R.T
    if method == "direct":
        lhs = np.eye(m * m) - np.kron(T, T)
        P0 = np.linalg.solve(lhs, RQR.reshape(-1)).reshape(m, m)
    else:
        P0 = scipy.linalg.solve_discrete_lyapunov(T, RQR, method="bilinear")
    return a0, P0


def stationary_init_torch(c, T, R, Q):
    """Batched (leading dims broadcast) differentiable version of the "direct" method."""
    m = T.shape[-1]
    eye = torch.eye(m, dtype=T.dtype)
    a0 = torch.linalg.solve(eye - T, c.unsqueeze(-1)).squeeze(-1)
    RQR = R @ Q @ R.transpose(-1, -2)
    kron = torch.einsum("...ij,...kl->...ikjl", T, T).reshape(*T.shape[:-2], m * m, m * m)
    lhs = torch.eye(m * m, dtype=T.dtype) - kron
    P0 = torch.linalg.solve(lhs, RQR.reshape(*RQR.shape[:-2], m * m, 1)).reshape(*RQR.shape[:-2], m, m)
    return a0, P0


def stationary_init_grad(c, T, R, Q, a0_bar, P0_bar):
    """Gradients of <a0_bar, a0> + <P0_bar, P0> w.r.t. (c, T, R, Q) by torch autograd."""
    as_t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
    prm = [as_t(x).clone().requires_grad_(True) for x in (c, T, R, Q)]
    a0, P0 = stationary_init_torch(*prm)
    obj = (a0 * as_t(a0_bar)).sum() + (P0 * as_t(P0_bar)).sum()
    gs = torch.autograd.grad(obj, prm)
    return a0.detach().numpy(), P0.detach().numpy(), {k: g.numpy() for k, g in zip(("c", "T", "R", "Q"), gs)}
