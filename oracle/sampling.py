"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the draw the reference makes from the smoother's output.

Reference: ``SequenceMvNormal.rv_op`` (pymc_extras/statespace/filters/distributions.py:411-443) scans
``MvNormalSVD.dist(mu=mu, cov=cov)`` (:52-87) over time; on the default backend that is NumPy's
``Generator.multivariate_normal(mean, cov, method="svd")``:  ``x = mean + z @ (sqrt(s)[:, None] * vh)``  with
``u, s, vh = svd(cov)`` and ``z`` standard normal.

``svd_factor`` restates that square root; ``chol_psd`` is the square root the CUDA kernel uses (lower Cholesky with
semidefinite pivots dropped, same tolerance).  Both satisfy ``A^T A = cov`` / ``L L^T = cov`` for a PSD matrix, so the
draws ``mean + z @ A`` and ``mean + L z`` have the same distribution; the tests pin the kernel to ``chol_psd`` exactly
and both factors to ``cov``.  Parity with the reference's NUMBERS is not defined for a random draw (its own stream is not
reproducible across generators); this is stated in DESIGN.md.
"""

import numpy as np


def svd_factor(cov):
    """A with x = mean + z @ A as numpy.random.Generator.multivariate_normal(method="svd") forms it."""
    u, s, vh = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * vh


def chol_psd(cov):
    """Lower factor L, L L^T = cov for PSD cov; a pivot <= 64 m eps max(diag) zeroes its column."""
    cov = np.asarray(cov, dtype=np.float64)
    m = cov.shape[-1]
    tol = 64 * m * np.finfo(np.float64).eps * np.max(np.diag(cov))
    L = np.zeros((m, m))
    for j in range(m):
        acc = cov[j:, j] - L[j:, :j] @ L[j, :j]
        d = acc[0]
        if d > tol:
            L[j, j] = np.sqrt(d)
            L[j + 1:, j] = acc[1:] / L[j, j]
    return L


def sample(mu, cov, z):
    """x[..., t, :] = mu + L z with L = chol_psd(cov[..., t, :, :]) (loops: small cases only)."""
    mu, cov, z = np.asarray(mu, float), np.asarray(cov, float), np.asarray(z, float)
    x = np.empty_like(mu)
    for idx in np.ndindex(mu.shape[:-1]):
        x[idx] = mu[idx] + chol_psd(cov[idx]) @ z[idx]
    return x


def simulate(a0, P0, c, d, T, Z, R, H, Q, z_x0, z_y0, z_state, z_obs, time_varying=()):
    """One model: restatement of LinearGaussianStateSpace.rv_op's step_fn / initial draw (distributions.py:239-263) with
    the normal draws written as chol_psd(cov) @ z.  ``time_varying``: names of the inputs that are sequences of length n
    (``sequence_names``; the initial observation uses Z[0], H[0], :259-260).  Returns (states (n+1, m), obs (n+1, p))."""
    vals = dict(c=c, d=d, T=T, Z=Z, R=R, H=H, Q=Q)
    at = lambda k, t: vals[k][t] if k in time_varying else vals[k]
    n = z_state.shape[0]
    xs = np.empty((n + 1, a0.shape[0]))
    ys = np.empty((n + 1, z_y0.shape[0]))
    xs[0] = a0 + chol_psd(P0) @ z_x0                                  # init_x_ = MvNormalSVD(a0, P0)
    ys[0] = at("Z", 0) @ xs[0] + chol_psd(at("H", 0)) @ z_y0          # init_y_ = MvNormalSVD(Z @ init_x_, H): no d
    for t in range(n):
        xs[t + 1] = at("c", t) + at("T", t) @ xs[t] + at("R", t) @ (chol_psd(at("Q", t)) @ z_state[t])
        ys[t + 1] = at("d", t) + at("Z", t) @ xs[t + 1] + chol_psd(at("H", t)) @ z_obs[t]
    return xs, ys
