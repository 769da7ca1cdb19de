"""CPU oracle: NumPy/SciPy float64 restatement of the pymc_extras.statespace Kalman filters.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker / the CPU baseline.  The product path (``pymc_experimental_b200``) never imports it.

Every function follows the reference source line by line (paths relative to ``/root/reference``):

* ``pymc_extras/statespace/filters/kalman_filter.py``   (BaseFilter / StandardFilter / SquareRootFilter /
  UnivariateFilter)
* ``pymc_extras/statespace/filters/kalman_smoother.py`` (KalmanSmoother)
* ``pymc_extras/statespace/filters/utilities.py:50-59`` (stabilize, quad_form_sym)
* ``pymc_extras/statespace/utils/constants.py:19-20``   (MISSING_FILL, JITTER_DEFAULT)

The arithmetic of the reference lives in PyTensor (third-party, not vendored, version unpinned:
``requirements.txt:1`` -> ``pymc>=5.19.1``); its linalg Ops call the same host LAPACK routines used here
(``scipy.linalg.solve(assume_a="pos")`` -> ``?posv``, ``numpy.linalg.det`` -> ``?getrf``,
``numpy.linalg.qr(mode="r")`` -> ``?geqrf``, ``numpy.linalg.pinv`` -> ``?gesdd``).

Pinning status (see DESIGN.md "Oracle"): this restatement is checked, in ``tests/test_oracle.py``, against
golden vectors produced by executing the reference's *own source files* through an eager PyTensor-API shim
(``oracle/ptshim``, script ``tests/golden/generate_golden.py``), against the reference's scipy-MVN identity
test (``tests/statespace/test_distributions.py:96-119``) and against a textbook (jitter-free) Kalman
filter/smoother standing in for statsmodels at the reference's own 1e-6 tolerance
(``tests/statespace/test_kalman_filter.py:232-269``).  Gradients are not pinned by any reference test.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla

# pymc_extras/statespace/utils/constants.py:19-20
MISSING_FILL = -9999.0
JITTER_DEFAULT = 1e-8
JITTER_DEFAULT_F32 = 1e-6
# pymc_extras/statespace/filters/kalman_filter.py:21
MVN_CONST = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------------------------
# filters/utilities.py:50-59
# --------------------------------------------------------------------------------------------------------
def stabilize(cov, jitter=JITTER_DEFAULT):
    """utilities.py:50-54 -- ``cov + I * jitter``."""
    return cov + np.eye(cov.shape[-1], dtype=cov.dtype) * jitter


def quad_form_sym(A, B):
    """utilities.py:57-59 -- ``out = A B A^T ; 0.5 (out + out^T)``."""
    out = A @ B @ A.T
    return 0.5 * (out + out.T)


def _at(x, t, static_ndim):
    """Time-varying inputs carry a leading time dimension (filters/utilities.py:8-29)."""
    x = np.asarray(x)
    return x[t] if x.ndim == static_ndim + 1 else x


# --------------------------------------------------------------------------------------------------------
# BaseFilter pieces (kalman_filter.py:310-410)
# --------------------------------------------------------------------------------------------------------
def handle_missing_values(y, Z, H, missing_fill_value):
    """kalman_filter.py:352-360."""
    nan_mask = np.isnan(y) | (y == missing_fill_value)
    all_nan_flag = bool(np.all(nan_mask))
    W = np.diag((~nan_mask).astype(y.dtype))
    Z_masked = W @ Z
    H_masked = W @ H
    y_masked = y.copy()
    y_masked[nan_mask] = 0.0
    return y_masked, Z_masked, H_masked, all_nan_flag


def predict(a, P, c, T, R, Q):
    """kalman_filter.py:407-410."""
    a_hat = T @ a + c
    P_hat = quad_form_sym(T, P) + quad_form_sym(R, Q)
    return a_hat, P_hat


# --------------------------------------------------------------------------------------------------------
# StandardFilter.update (kalman_filter.py:594-617)
# --------------------------------------------------------------------------------------------------------
def standard_update(a, P, y, d, Z, H, all_nan_flag, jitter):
    m = a.shape[0]
    y_hat = d + Z @ a
    v = y - y_hat

    PZT = P @ Z.T
    F = Z @ PZT + stabilize(H, jitter)

    K = sla.solve(F.T, PZT.T, assume_a="pos", check_finite=False).T
    I_KZ = np.eye(m, dtype=a.dtype) - K @ Z

    a_filtered = a + K @ v
    P_filtered = quad_form_sym(I_KZ, P) + quad_form_sym(K, H)

    F_inv_v = sla.solve(F, v, assume_a="pos", check_finite=False)
    inner_term = v @ F_inv_v
    F_logdet = np.log(np.linalg.det(F))

    ll = 0.0 if all_nan_flag else -0.5 * (MVN_CONST + F_logdet + inner_term)
    return a_filtered, P_filtered, y_hat, F, ll


# --------------------------------------------------------------------------------------------------------
# SquareRootFilter (kalman_filter.py:629-716)
# --------------------------------------------------------------------------------------------------------
def sqrt_predict(a, P_chol, c, T, R, Q):
    """kalman_filter.py:639-648."""
    m = a.shape[0]
    a_hat = T @ a + c
    Q_chol = np.linalg.cholesky(Q)
    M = np.hstack([T @ P_chol, R @ Q_chol]).T
    R_decomp = np.linalg.qr(M, mode="r")
    P_chol_hat = R_decomp[:m, :m].T
    return a_hat, P_chol_hat


def sqrt_update(a, P_chol, y, d, Z, H, all_nan_flag):
    """kalman_filter.py:661-716."""
    m = a.shape[0]
    p = y.shape[0]
    y_hat = Z @ a + d
    v = y - y_hat

    H_chol = H if np.all(H == 0.0) else np.linalg.cholesky(H)

    zeros = np.zeros((m, p), dtype=a.dtype)
    upper = np.hstack([H_chol, Z @ P_chol])
    lower = np.hstack([zeros, P_chol])
    A_T = np.vstack([upper, lower])
    B = np.linalg.qr(A_T.T, mode="r").T

    F_chol = B[:p, :p]
    K_F_chol = B[p:, :p]
    P_chol_filtered = B[p:, p:]

    if all_nan_flag:
        # compute_degenerate, :700-705
        return a, P_chol, y_hat, F_chol, 0.0

    s = sla.solve_triangular(F_chol, v, lower=True, check_finite=False)
    a_filtered = a + K_F_chol @ s
    inner_term = sla.solve_triangular(F_chol, s, lower=True, check_finite=False)
    loss = v @ inner_term
    logdet = 2.0 * np.log(np.abs(np.diag(F_chol))).sum()
    ll = -0.5 * (p * (MVN_CONST + logdet) + loss)
    return a_filtered, P_chol_filtered, y_hat, F_chol, ll


# --------------------------------------------------------------------------------------------------------
# UnivariateFilter (kalman_filter.py:769-820)
# --------------------------------------------------------------------------------------------------------
def univariate_inner_step(y, Z_row, d_row, sigma_H, nan_flag, a, P, jitter):
    """kalman_filter.py:769-789."""
    y_hat = Z_row @ a + d_row
    v = y - y_hat

    PZT = P @ Z_row
    F = Z_row @ PZT + sigma_H

    F_zero_flag = (F == 0) or bool(nan_flag)
    F = F + jitter

    K = PZT / F * (1.0 - float(F_zero_flag))
    a_filtered = a + K * v
    P_filtered = P - np.outer(K, K) * F
    ll_inner = 0.0 if F_zero_flag else np.log(F) + v**2 / F
    return a_filtered, P_filtered, y_hat, F, ll_inner


def univariate_step(y, a, P, c, d, T, Z, R, H, Q, jitter):
    """kalman_filter.py:791-820 (note: isnan only, the fill value is ignored -- SURVEY quirk Q6)."""
    p = y.shape[0]
    nan_mask = np.isnan(y)
    W = np.eye(p, dtype=a.dtype)
    W[nan_mask, nan_mask] = 0.0
    Z_masked = W @ Z
    H_masked = W @ H
    y_masked = y.copy()
    y_masked[nan_mask] = 0.0
    H_diag = np.diag(H_masked)

    ll_inner = np.zeros(p, dtype=a.dtype)
    obs_mu = np.zeros(1, dtype=a.dtype)
    obs_cov = np.zeros((1, 1), dtype=a.dtype)
    a_f, P_f = a, P
    for i in range(p):
        a_f, P_f, y_hat, F, ll_inner[i] = univariate_inner_step(
            y_masked[i], Z_masked[i], d[i], H_diag[i], nan_mask[i], a_f, P_f, jitter
        )
        obs_mu = np.atleast_1d(y_hat)
        obs_cov = np.atleast_2d(F)

    P_f = stabilize(0.5 * (P_f + P_f.T), jitter)
    a_hat, P_hat = predict(a_f, P_f, c, T, R, Q)
    ll = -0.5 * (np.count_nonzero(ll_inner) * MVN_CONST + ll_inner.sum())
    return a_f, a_hat, obs_mu, P_f, P_hat, obs_cov, ll


# --------------------------------------------------------------------------------------------------------
# drivers == BaseFilter.build_graph + _postprocess_scan_results (kalman_filter.py:144-308, 718-750)
# --------------------------------------------------------------------------------------------------------
FILTER_KINDS = ("standard", "univariate", "cholesky")


def kalman_filter(kind, data, a0, P0, c, d, T, Z, R, H, Q, missing_fill_value=None, cov_jitter=None):
    """Run one filter over one series.  Returns the 7 outputs of ``BaseFilter.build_graph`` in order:
    filtered_states (n,m), predicted_states (n,m), observed_states (n,p), filtered_covariances (n,m,m),
    predicted_covariances (n,m,m), observed_covariances (n,p,p), loglike_obs (n,).

    For the univariate filter observed_states/observed_covariances have trailing dims 1 / (1,1)
    (last inner iteration only, kalman_filter.py:808-813).
    """
    if missing_fill_value is None:
        missing_fill_value = MISSING_FILL
    if cov_jitter is None:
        cov_jitter = JITTER_DEFAULT
    data = np.asarray(data, dtype=np.float64)
    a0, P0 = np.asarray(a0, dtype=np.float64), np.asarray(P0, dtype=np.float64)
    n, p = data.shape
    m = np.asarray(R).shape[-2]

    a, P = a0, P0
    outs = [[] for _ in range(7)]
    for t in range(n):
        y = data[t]
        ct, dt = _at(c, t, 1), _at(d, t, 1)
        Tt, Zt, Rt, Ht, Qt = _at(T, t, 2), _at(Z, t, 2), _at(R, t, 2), _at(H, t, 2), _at(Q, t, 2)
        if kind == "univariate":
            a_f, a_hat, obs_mu, P_f, P_hat, obs_cov, ll = univariate_step(
                y, a, P, ct, dt, Tt, Zt, Rt, Ht, Qt, cov_jitter
            )
        else:
            y_m, Z_m, H_m, flag = handle_missing_values(y, Zt, Ht, missing_fill_value)
            if kind == "standard":
                a_f, P_f, obs_mu, obs_cov, ll = standard_update(a, P, y_m, dt, Z_m, H_m, flag, cov_jitter)
                P_f = stabilize(P_f, cov_jitter)
                a_hat, P_hat = predict(a_f, P_f, ct, Tt, Rt, Qt)
            elif kind == "cholesky":
                a_f, P_f, obs_mu, obs_cov, ll = sqrt_update(a, P, y_m, dt, Z_m, H_m, flag)
                P_f = stabilize(P_f, cov_jitter)  # jitter on the *factor*, kalman_filter.py:536 (quirk Q4)
                a_hat, P_hat = sqrt_predict(a_f, P_f, ct, Tt, Rt, Qt)
            else:
                raise ValueError(kind)
        for lst, val in zip(outs, (a_f, a_hat, obs_mu, P_f, P_hat, obs_cov, ll)):
            lst.append(val)
        a, P = a_hat, P_hat

    filt_a, pred_a, obs_mu, filt_P, pred_P, obs_cov, ll = (np.array(o) for o in outs)
    # _postprocess_scan_results, :264-269
    pred_a = np.concatenate([a0[None], pred_a[:-1]], axis=0)
    pred_P = np.concatenate([P0[None], pred_P[:-1]], axis=0)
    if kind == "cholesky":
        # :733-740
        sq = lambda L: np.einsum("...ij,...kj->...ik", L, L)
        filt_P, pred_P, obs_cov = sq(filt_P), sq(pred_P), sq(obs_cov)
    return filt_a, pred_a, obs_mu, filt_P, pred_P, obs_cov, ll


def kalman_logp(kind, data, a0, P0, c, d, T, Z, R, H, Q, **kw):
    return float(kalman_filter(kind, data, a0, P0, c, d, T, Z, R, H, Q, **kw)[-1].sum())


# --------------------------------------------------------------------------------------------------------
# KalmanSmoother (kalman_smoother.py:66-126)
# --------------------------------------------------------------------------------------------------------
def kalman_smoother(T, R, Q, filtered_states, filtered_covariances, cov_jitter=JITTER_DEFAULT):
    n, m = filtered_states.shape
    a_s = filtered_states[-1]
    P_s = filtered_covariances[-1]
    sm_a = np.empty_like(filtered_states)
    sm_P = np.empty_like(filtered_covariances)
    sm_a[-1], sm_P[-1] = a_s, P_s
    for t in range(n - 2, -1, -1):
        a, P = filtered_states[t], filtered_covariances[t]
        # time-varying T/R/Q: scan(go_backwards=True) reverses each sequence and then truncates to the shortest
        # (filtered[:-1] has n-1 rows), so filtered[t] meets x[t+1] (kalman_smoother.py:84-92; DESIGN quirk Q9)
        Tt, Rt, Qt = _at(T, t + 1, 2), _at(R, t + 1, 2), _at(Q, t + 1, 2)
        # predict, :121-126 (no state intercept, quirk Q8)
        a_hat = Tt @ a
        P_hat = stabilize(quad_form_sym(Tt, P) + quad_form_sym(Rt, Qt), cov_jitter)
        # smoother_step, :111-117
        gain = (np.linalg.pinv(P_hat) @ Tt @ P).T
        a_s = a + gain @ (a_s - a_hat)
        P_next = P + quad_form_sym(gain, P_s - P_hat)
        P_next = stabilize(P_next, cov_jitter)
        P_s = stabilize(P_next)  # second jitter is always JITTER_DEFAULT (:117)
        sm_a[t], sm_P[t] = a_s, P_s
    return sm_a, sm_P
