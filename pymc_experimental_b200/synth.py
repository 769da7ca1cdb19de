"""Seeded synthetic state-space inputs for the BASELINE.json configs.

Matrix layouts follow the reference model zoo (paths relative to /root/reference/pymc_extras/statespace):
  C1  models/structural.py:858-875      LevelTrend(order=2)                       (m,p,r) = (2,1,2)
  C2  models/SARIMAX.py:383-536         SARIMA(1,1,1)x(1,0,1,12), Harvey form     (15,1,1)
  C3  models/structural.py:503-564      trend + TimeSeasonality(12) + Cycle       (15,1,5)
  C4  models/VARMAX.py:297-393          VARMAX(2,0), k_endog=16                   (32,16,16)
  C5  models/ETS.py:468-670             ETS(A,Ad,N)                               (3,1,1)

Every function returns a dict with float64 C-contiguous arrays
  y (T,p) [shared] or (B,T,p);  a0 (B,m); P0 (B,m,m); c (B,m); d (B,p); T (B,m,m); Z (B,p,m); R (B,m,r);
  H (B,p,p); Q (B,r,r)
ready for :func:`pymc_experimental_b200.engine.KalmanEngine.logp_grad`.
"""

from __future__ import annotations

import numpy as np

SEED_BASE = 1234


def _simulate(rng, T_, Z, R, H, Q, a0, n):
    """Simulate y_t = Z x_t + eps, x_{t+1} = T x_t + R eta from one parameter set."""
    m, r = R.shape
    p = Z.shape[0]
    Lq = np.linalg.cholesky(Q + 1e-12 * np.eye(r))
    Lh = np.linalg.cholesky(H + 1e-12 * np.eye(p))
    x = a0.copy()
    y = np.empty((n, p))
    for t in range(n):
        y[t] = Z @ x + Lh @ rng.standard_normal(p)
        x = T_ @ x + R @ (Lq @ rng.standard_normal(r))
    return y


def _bcast(x, B):
    return np.ascontiguousarray(np.broadcast_to(x, (B, *x.shape)))


def local_level(B=1, n=100, seed=SEED_BASE + 1):
    """C1: local linear trend, Nile-like (tests/statespace/utilities/test_helpers.py:138-170)."""
    rng = np.random.default_rng(seed)
    T_ = np.array([[1.0, 1.0], [0.0, 1.0]])
    Z = np.array([[1.0, 0.0]])
    R = np.eye(2)
    Q = np.diag([0.5, 0.01])
    H = np.array([[0.8]])
    a0 = np.zeros(2)
    P0 = np.eye(2) * 1e6
    y = _simulate(rng, T_, Z, R, H, Q, a0, n)
    return dict(y=y, a0=_bcast(a0, B), P0=_bcast(P0, B), c=np.zeros((B, 2)), d=np.zeros((B, 1)),
                T=_bcast(T_, B), Z=_bcast(Z, B), R=_bcast(R, B), H=_bcast(H, B), Q=_bcast(Q, B))


def sarima(B=16384, n=500, seed=SEED_BASE + 2):
    """C2: SARIMA(1,1,1)x(1,0,1,12), m=15.  One shared series, B parameter draws."""
    rng = np.random.default_rng(seed)
    m = 15
    phi = rng.uniform(-0.8, 0.8, B)
    Phi = rng.uniform(-0.8, 0.8, B)
    theta = rng.uniform(-0.5, 0.5, B)
    Theta = rng.uniform(-0.5, 0.5, B)
    sigma = np.exp(rng.normal(0.0, 0.3, B))
    T_ = np.zeros((B, m, m))
    T_[:, 0, 0] = 1.0
    T_[:, 0, 1] = 1.0
    for i in range(1, 14):
        T_[:, i, i + 1] = 1.0
    T_[:, 1, 1] = phi
    T_[:, 12, 1] = Phi
    T_[:, 13, 1] = -Phi * phi
    R = np.zeros((B, m, 1))
    R[:, 1, 0] = 1.0
    R[:, 2, 0] = theta
    R[:, 13, 0] = Theta
    R[:, 14, 0] = Theta * theta
    Z = np.zeros((B, 1, m))
    Z[:, 0, 0] = 1.0
    Z[:, 0, 1] = 1.0
    Q = (sigma**2).reshape(B, 1, 1)
    H = np.zeros((B, 1, 1))
    a0 = np.zeros((B, m))
    P0 = _bcast(10.0 * np.eye(m), B)
    y = _simulate(rng, T_[0], Z[0], R[0], H[0], Q[0], a0[0], n)
    return dict(y=y, a0=a0, P0=P0, c=np.zeros((B, m)), d=np.zeros((B, 1)), T=T_, Z=Z, R=R, H=H,
                Q=np.ascontiguousarray(Q))


# (matrix, index) of every entry of the C2 model that depends on the draw (models/SARIMAX.py:383-536): phi, Phi and
# -Phi*phi in the transition matrix, theta, Theta and Theta*theta in the selection matrix, sigma^2 in the state covariance
SARIMA_POSITIONS = [("T", (1, 1)), ("T", (12, 1)), ("T", (13, 1)), ("R", (2, 0)), ("R", (13, 0)), ("R", (14, 0)),
                    ("Q", (0, 0))]


def sarima_compact(cfg):
    """(base, positions, u) of a ``sarima`` config for ``KalmanEngine.logp_grad_compact``."""
    names = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
    base = {k: np.array(cfg[k][0]) for k in names}
    u = np.stack([cfg[name][(slice(None), *idx)] for name, idx in SARIMA_POSITIONS], axis=1)
    return base, SARIMA_POSITIONS, np.ascontiguousarray(u)


def structural(B=65536, n=240, seed=SEED_BASE + 3):
    """C3: LevelTrend(2) + TimeSeasonality(12) + Cycle, m=15, r=5, with MeasurementError."""
    rng = np.random.default_rng(seed)
    m, r = 15, 5
    lam = 2.0 * np.pi / 40.0
    rho = 0.95
    T1 = np.zeros((m, m))
    T1[0, 0] = T1[0, 1] = T1[1, 1] = 1.0
    T1[2, 2:13] = -1.0
    for i in range(3, 13):
        T1[i, i - 1] = 1.0
    T1[13:, 13:] = rho * np.array([[np.cos(lam), np.sin(lam)], [-np.sin(lam), np.cos(lam)]])
    Z1 = np.zeros((1, m))
    Z1[0, 0] = Z1[0, 2] = Z1[0, 13] = 1.0
    R1 = np.zeros((m, r))
    R1[0, 0] = R1[1, 1] = R1[2, 2] = R1[13, 3] = R1[14, 4] = 1.0
    sig = np.exp(rng.normal(-1.0, 0.3, (B, 4)))  # level, trend, season, cycle
    Q = np.zeros((B, r, r))
    Q[:, 0, 0] = sig[:, 0] ** 2
    Q[:, 1, 1] = sig[:, 1] ** 2
    Q[:, 2, 2] = sig[:, 2] ** 2
    Q[:, 3, 3] = Q[:, 4, 4] = sig[:, 3] ** 2
    H = (np.exp(rng.normal(-1.0, 0.3, B)) ** 2).reshape(B, 1, 1)
    a0 = rng.standard_normal((B, m))
    P0 = _bcast(np.eye(m), B)
    y = _simulate(rng, T1, Z1, R1, H[0], Q[0], a0[0], n)
    return dict(y=y, a0=a0, P0=P0, c=np.zeros((B, m)), d=np.zeros((B, 1)), T=_bcast(T1, B), Z=_bcast(Z1, B),
                R=_bcast(R1, B), H=np.ascontiguousarray(H), Q=Q)


def varmax(B=4096, n=1000, k=16, seed=SEED_BASE + 4):
    """C4: VARMAX(2,0), k_endog=k, m=2k; P0 is a Cholesky-like factor."""
    rng = np.random.default_rng(seed)
    m = 2 * k
    A = rng.normal(0.0, 0.05, (B, k, m))
    T_ = np.zeros((B, m, m))
    T_[:, :k, :] = A
    T_[:, k:, :k] = np.eye(k)
    # rescale to spectral radius <= 0.9 (cheap bound: infinity norm of the AR block)
    nrm = np.abs(A).sum(-1).max(-1)
    T_[:, :k, :] *= np.minimum(1.0, 0.9 / nrm)[:, None, None]
    R = np.zeros((m, k))
    R[:k] = np.eye(k)
    Z = np.zeros((k, m))
    Z[:, :k] = np.eye(k)
    L = np.tril(rng.normal(0.0, 0.1, (B, k, k)), -1) + 0.5 * np.eye(k)
    Q = L @ L.transpose(0, 2, 1)
    H = np.zeros((B, k, k))
    idx = np.arange(k)
    H[:, idx, idx] = rng.uniform(0.05, 0.2, (B, k))
    a0 = np.zeros((B, m))
    P0 = _bcast(np.eye(m), B)
    y = _simulate(rng, T_[0], Z, R, H[0], Q[0], a0[0], n)
    return dict(y=y, a0=a0, P0=P0, c=np.zeros((B, m)), d=np.zeros((B, k)), T=T_, Z=_bcast(Z, B), R=_bcast(R, B),
                H=H, Q=Q)


def ets(B=1_000_000, n=100, seed=SEED_BASE + 5):
    """C5: ETS(A,Ad,N): states [innovation, level, trend]; every series has its own y."""
    rng = np.random.default_rng(seed)
    alpha = rng.uniform(0.1, 0.9, B)
    beta = rng.uniform(0.05, 0.5, B)
    phi = rng.uniform(0.8, 0.98, B)
    sigma = np.exp(rng.normal(0.0, 0.3, B))
    T_ = np.zeros((B, 3, 3))
    T_[:, 1, 1] = 1.0
    T_[:, 1, 2] = phi
    T_[:, 2, 2] = phi
    R = np.stack([1.0 - alpha, alpha, alpha * beta], axis=-1)[:, :, None]
    Z = _bcast(np.array([[1.0, 1.0, 0.0]]), B)
    Q = (sigma**2).reshape(B, 1, 1)
    H = np.zeros((B, 1, 1))
    a0 = np.zeros((B, 3))
    a0[:, 1:] = rng.standard_normal((B, 2))
    P0 = _bcast(np.eye(3), B)
    # vectorised simulation: one series per parameter set
    x0, x1, x2 = a0[:, 0].copy(), a0[:, 1].copy(), a0[:, 2].copy()
    Y = np.empty((n, B))
    E = rng.standard_normal((n, B))
    r0, r1, r2 = R[:, 0, 0] * sigma, R[:, 1, 0] * sigma, R[:, 2, 0] * sigma
    for t in range(n):  # x <- T x + R e with T = [[0,0,0],[0,1,phi],[0,0,phi]] written out (a million 3 x 3 einsums are slow)
        Y[t] = x0 + x1
        e = E[t]
        x0, x1, x2 = r0 * e, x1 + phi * x2 + r1 * e, phi * x2 + r2 * e
    y = np.ascontiguousarray(Y.T)[:, :, None]
    return dict(y=y, a0=a0, P0=P0, c=np.zeros((B, 3)), d=np.zeros((B, 1)), T=T_, Z=Z,
                R=np.ascontiguousarray(R), H=H, Q=np.ascontiguousarray(Q))


def random_dense(B, n, m, p, r, seed=0, missing=0.0, per_series_y=True, diag_H=False):
    """Dense random SPD H, Q, P0 and dense T, Z, R, non-zero c, d: exercises every term of the adjoint."""
    rng = np.random.default_rng(seed)

    def spd(k, scale):
        A = rng.standard_normal((B, k, k)) * scale
        return A @ A.transpose(0, 2, 1) + 0.5 * scale * np.eye(k)

    T_ = rng.standard_normal((B, m, m)) * (0.6 / np.sqrt(m))
    Z = rng.standard_normal((B, p, m))
    R = rng.standard_normal((B, m, r))
    y = rng.standard_normal((B, n, p) if per_series_y else (n, p)) * 2.0
    if missing > 0:
        y[rng.uniform(size=y.shape) < missing] = np.nan
    H = spd(p, 0.7)
    if diag_H:  # keeps F symmetric when some (not all) observations of a step are missing
        H = H * np.eye(p)
    return dict(y=y, a0=rng.standard_normal((B, m)), P0=spd(m, 1.0), c=rng.standard_normal((B, m)) * 0.1,
                d=rng.standard_normal((B, p)) * 0.1, T=T_, Z=Z, R=R, H=H, Q=spd(r, 0.5))


def random_dense_tv(B, n, m, p, r, time_varying, seed=0, missing=0.0, diag_H=False):
    """``random_dense`` with a time axis ([B, n, ...]) on the inputs named in ``time_varying`` (c, d, T, Z, R, H, Q):
    every step gets its own perturbed copy, H_t and Q_t stay SPD.  Mirrors the reference's time-varying tests
    (tests/statespace/test_kalman_filter.py:151-166) with non-trivial values."""
    cfg = random_dense(B, n, m, p, r, seed=seed, missing=missing, diag_H=diag_H)
    rng = np.random.default_rng(seed + 7919)
    for k in time_varying:
        x = cfg[k]
        xt = np.repeat(x[:, None], n, axis=1)
        if k in ("H", "Q"):
            kdim = x.shape[-1]
            L = rng.standard_normal((B, n, kdim, kdim)) * 0.25
            pert = L @ np.swapaxes(L, -1, -2)
            if k == "H" and diag_H:
                pert = pert * np.eye(kdim)
            xt = xt + pert
        else:
            scale = 0.15 * (np.abs(x).mean() + 0.05)
            xt = xt + rng.standard_normal(xt.shape) * scale
        cfg[k] = np.ascontiguousarray(xt)
    return cfg


CONFIGS = {"C1": local_level, "C2": sarima, "C3": structural, "C4": varmax, "C5": ets}
FILTER_OF = {"C1": "standard", "C2": "univariate", "C3": "standard", "C4": "cholesky", "C5": "standard"}
