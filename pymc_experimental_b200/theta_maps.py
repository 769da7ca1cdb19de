"""theta -> matrix-entry programs for ``KalmanEngine.logp_grad_theta`` (``bkf_filter_logp_grad_theta``, SURVEY section 8 row
f-1): the free parameters of the reference's model families written as the data the C ABI takes -- positions
``(matrix, index)`` of the entries that depend on the draw, and per entry a sum of terms ``coef * prod fn(theta[i])``.
No model code enters the library; these builders are the host-side mirror of the reference's ``make_symbolic_graph``
methods, restricted to the part of each model that depends on the parameters.

``evaluate`` / ``jacobian`` restate the program in NumPy (the checker of the device path, and a caller's way to inspect
what it ships).
"""

from __future__ import annotations

import numpy as np

from .engine import FN_COS, FN_EXP, FN_ID, FN_ONE_MINUS, FN_SIN, FN_SQUARE

_FN = {
    FN_ID: (lambda x: x, lambda x: np.ones_like(x)),
    FN_COS: (np.cos, lambda x: -np.sin(x)),
    FN_SIN: (np.sin, np.cos),
    FN_EXP: (np.exp, np.exp),
    FN_ONE_MINUS: (lambda x: 1.0 - x, lambda x: -np.ones_like(x)),
    FN_SQUARE: (lambda x: x * x, lambda x: 2.0 * x),
}


def evaluate(terms, theta, nnz):
    """u [B, nnz] of a program at theta [B, ntheta]."""
    theta = np.asarray(theta, dtype=np.float64)
    u = np.zeros((theta.shape[0], nnz))
    for k, coef, factors in terms:
        v = np.full(theta.shape[0], float(coef))
        for ti, fn in factors:
            v = v * _FN[fn][0](theta[:, ti])
        u[:, k] += v
    return u


def jacobian(terms, theta, nnz):
    """du/dtheta [B, nnz, ntheta] of a program."""
    theta = np.asarray(theta, dtype=np.float64)
    J = np.zeros((theta.shape[0], nnz, theta.shape[1]))
    for k, coef, factors in terms:
        for f, (ti, fn) in enumerate(factors):
            v = float(coef) * _FN[fn][1](theta[:, ti])
            for f2, (tj, fn2) in enumerate(factors):
                if f2 != f:
                    v = v * _FN[fn2][0](theta[:, tj])
            J[:, k, ti] += v
    return J


def sarima_111_101_12():
    """SARIMA(1,1,1)x(1,0,1,12) in the reference's Harvey representation (models/SARIMAX.py:404-458; BASELINE config C2).
    theta = (phi, Phi, theta_ma, Theta_ma, sigma): the multiplied-out lag polynomials put phi, Phi, -phi Phi into the
    transition matrix, theta, Theta, theta Theta into the selection matrix and sigma^2 into the state covariance.
    Returns ``(positions, terms, ntheta)``."""
    positions = [("T", (1, 1)), ("T", (12, 1)), ("T", (13, 1)), ("R", (2, 0)), ("R", (13, 0)), ("R", (14, 0)),
                 ("Q", (0, 0))]
    terms = [
        (0, 1.0, [(0, FN_ID)]),
        (1, 1.0, [(1, FN_ID)]),
        (2, -1.0, [(0, FN_ID), (1, FN_ID)]),
        (3, 1.0, [(2, FN_ID)]),
        (4, 1.0, [(3, FN_ID)]),
        (5, 1.0, [(2, FN_ID), (3, FN_ID)]),
        (6, 1.0, [(4, FN_SQUARE)]),
    ]
    return positions, terms, 5


def structural_cycle(first_state, shock, log_sigma=True):
    """Stochastic cycle block of the structural models (models/structural.py:1220-1237): T[s:s+2, s:s+2] =
    rho [[cos l, sin l], [-sin l, cos l]], Q[j, j] = Q[j+1, j+1] = sigma^2.  theta = (rho, lambda, log sigma | sigma).
    ``first_state``: index s of the block's first state, ``shock``: index j of its first shock."""
    s, j = first_state, shock
    positions = [("T", (s, s)), ("T", (s, s + 1)), ("T", (s + 1, s)), ("T", (s + 1, s + 1)), ("Q", (j, j)),
                 ("Q", (j + 1, j + 1))]
    sig = [(2, FN_EXP), (2, FN_EXP)] if log_sigma else [(2, FN_SQUARE)]
    terms = [
        (0, 1.0, [(0, FN_ID), (1, FN_COS)]),
        (1, 1.0, [(0, FN_ID), (1, FN_SIN)]),
        (2, -1.0, [(0, FN_ID), (1, FN_SIN)]),
        (3, 1.0, [(0, FN_ID), (1, FN_COS)]),
        (4, 1.0, sig),
        (5, 1.0, sig),
    ]
    return positions, terms, 3


def ets_aan(damped=True):
    """ETS(A,Ad,N) in the reference's innovations form (models/ETS.py:499-555): states (level, trend) after the error
    state; theta = (alpha, beta_star, phi, sigma) with beta = alpha beta_star.  T = [[0,0,0],[0,1,phi],[0,0,phi]],
    R = [1, alpha, alpha beta_star]^T, Q = sigma^2."""
    positions = [("R", (1, 0)), ("R", (2, 0)), ("T", (1, 2)), ("T", (2, 2)), ("Q", (0, 0))]
    terms = [
        (0, 1.0, [(0, FN_ID)]),
        (1, 1.0, [(0, FN_ID), (1, FN_ID)]),
        (4, 1.0, [(3, FN_SQUARE)]),
    ]
    if damped:
        terms += [(2, 1.0, [(2, FN_ID)]), (3, 1.0, [(2, FN_ID)])]
    else:
        terms += [(2, 1.0, []), (3, 1.0, [])]
    return positions, terms, 4
