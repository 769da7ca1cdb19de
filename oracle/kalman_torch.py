"""CPU oracle, differentiable: torch.float64 restatement of the reference Kalman filters (batched).

TEST INFRASTRUCTURE ONLY (see ``oracle/kalman_np.py`` header for the rules and the reference file:line map).

Same formulas as ``oracle/kalman_np.py`` -- i.e. ``pymc_extras/statespace/filters/kalman_filter.py`` --
written with torch ops so that ``torch.autograd`` yields d(sum_t ll_t)/d{a0,P0,c,d,T,Z,R,H,Q}: the quantity
``pytensor.grad`` derives through ``Scan.L_op`` for NUTS.  Every input may carry
arbitrary leading batch dimensions (broadcast against each other), which makes this the gradient oracle for
the batched CUDA engine and a multi-threaded CPU baseline.  ``time_varying`` names the inputs that carry a
time axis right before their static dims (filters/utilities.py:8-47).

Gradient conventions reproduced here (what autodiff of the reference formulas yields):
* every matrix entry is an independent variable (no symmetrisation of dH, dQ, dP0);
* ``solve(assume_a="pos")`` / ``det`` use the generic (un-symmetrised) adjoints (PyTensor ``SolveBase.L_op``,
  ``Det.grad``), identical to torch's ``linalg.solve`` / ``linalg.det`` adjoints.
"""

from __future__ import annotations

import math

import torch

MISSING_FILL = -9999.0
JITTER_DEFAULT = 1e-8
MVN_CONST = math.log(2.0 * math.pi)


def _eye_like(X):
    return torch.eye(X.shape[-1], dtype=X.dtype, device=X.device)


def _mT(X):
    return X.transpose(-1, -2)


def _mv(A, x):
    return (A @ x.unsqueeze(-1)).squeeze(-1)


def quad_form_sym(A, B):
    out = A @ B @ _mT(A)  # utilities.py:57-59
    return 0.5 * (out + _mT(out))


def predict(a, P, c, T, R, Q):
    return _mv(T, a) + c, quad_form_sym(T, P) + quad_form_sym(R, Q)  # kalman_filter.py:407-408


def _mask(y, Z, H, nan_mask):
    """kalman_filter.py:352-360 with W = diag(~mask): rows of Z and H are zeroed, y[mask] = 0."""
    w = (~nan_mask).to(Z.dtype)
    Zm = w.unsqueeze(-1) * Z
    Hm = w.unsqueeze(-1) * H
    ym = torch.where(nan_mask, torch.zeros_like(y), y)
    return ym, Zm, Hm, nan_mask.all(dim=-1)


def standard_step(y, a, P, c, d, T, Z, R, H, Q, fill, jitter):
    nan_mask = torch.isnan(y) | (y == fill)
    ym, Zm, Hm, flag = _mask(y, Z, H, nan_mask)
    eye_p = _eye_like(Hm)
    eye_m = _eye_like(P)
    y_hat = d + _mv(Zm, a)  # :594
    v = ym - y_hat
    PZT = P @ _mT(Zm)
    F = Zm @ PZT + (Hm + eye_p * jitter)  # :598
    K = _mT(torch.linalg.solve(_mT(F), _mT(PZT)))  # :600
    I_KZ = eye_m - K @ Zm
    a_f = a + _mv(K, v)
    P_f = quad_form_sym(I_KZ, P) + quad_form_sym(K, Hm)  # :604
    F_inv_v = torch.linalg.solve(F, v.unsqueeze(-1)).squeeze(-1)  # :606
    inner = (v * F_inv_v).sum(-1)
    logdet = torch.log(torch.linalg.det(F))  # :609
    ll = torch.where(flag, torch.zeros_like(inner), -0.5 * (MVN_CONST + logdet + inner))  # :611-615
    P_f = P_f + eye_m * jitter  # :536
    a_hat, P_hat = predict(a_f, P_f, c, T, R, Q)
    return a_f, a_hat, y_hat, P_f, P_hat, F, ll


def univariate_step(y, a, P, c, d, T, Z, R, H, Q, fill, jitter):
    """kalman_filter.py:769-820 (isnan only -- the fill value is ignored)."""
    nan_mask = torch.isnan(y)
    ym, Zm, Hm, _ = _mask(y, Z, H, nan_mask)
    p = y.shape[-1]
    Hd = torch.diagonal(Hm, dim1=-2, dim2=-1)
    ll_sum = 0.0
    n_nonzero = 0.0
    y_hat = F = None
    for i in range(p):
        z = Zm[..., i, :]
        y_hat = (z * a).sum(-1) + d[..., i]
        v = ym[..., i] - y_hat
        PZT = _mv(P, z)
        F0 = (z * PZT).sum(-1) + Hd[..., i]
        zero_flag = (F0 == 0) | nan_mask[..., i]
        F = F0 + jitter
        keep = (~zero_flag).to(a.dtype)
        K = PZT / F.unsqueeze(-1) * keep.unsqueeze(-1)
        a = a + K * v.unsqueeze(-1)
        P = P - K.unsqueeze(-1) * K.unsqueeze(-2) * F.unsqueeze(-1).unsqueeze(-1)
        ll_i = torch.where(zero_flag, torch.zeros_like(F), torch.log(F) + v**2 / F)
        ll_sum = ll_sum + ll_i
        n_nonzero = n_nonzero + (ll_i != 0).to(a.dtype)
    P_f = 0.5 * (P + _mT(P)) + _eye_like(P) * jitter  # :815
    a_hat, P_hat = predict(a, P_f, c, T, R, Q)
    ll = -0.5 * (n_nonzero * MVN_CONST + ll_sum)  # :818
    return a, a_hat, y_hat.unsqueeze(-1), P_f, P_hat, F.unsqueeze(-1).unsqueeze(-1), ll


def _qr_r(A):
    return torch.linalg.qr(A, mode="reduced")[1]


def _chol_lower(X):
    """LAPACK ?potrf(lower) reads only the lower triangle; differentiating through the explicit
    lower->symmetric fill reproduces PyTensor's ``Cholesky.L_op`` convention
    (``tril(s + s.T) - diag(s)``: the gradient lives in the lower triangle, off-diagonals doubled)."""
    Xl = torch.tril(X) + _mT(torch.tril(X, -1))
    return torch.linalg.cholesky(Xl)


def _safe_chol(X, degenerate):
    """cholesky(X) where ``degenerate`` is False; X itself where True (ifelse at kalman_filter.py:666)."""
    Xs = torch.where(degenerate[..., None, None], _eye_like(X).expand_as(X), X)
    L = _chol_lower(Xs)
    return torch.where(degenerate[..., None, None], X, L)


def sqrt_step(y, a, Pc, c, d, T, Z, R, H, Q, fill, jitter):
    """kalman_filter.py:629-716; Pc is the carried Cholesky-like factor."""
    nan_mask = torch.isnan(y) | (y == fill)
    ym, Zm, Hm, flag = _mask(y, Z, H, nan_mask)
    m, p = a.shape[-1], y.shape[-1]
    y_hat = _mv(Zm, a) + d
    v = ym - y_hat
    H_zero = (Hm == 0).all(-1).all(-1)
    Hc = _safe_chol(Hm, H_zero)
    ZP = Zm @ Pc
    bshape = torch.broadcast_shapes(Hc.shape[:-2], ZP.shape[:-2], Pc.shape[:-2])
    upper = torch.cat([Hc.expand(*bshape, p, p), ZP.expand(*bshape, p, m)], dim=-1)
    lower = torch.cat([torch.zeros(*bshape, m, p, dtype=a.dtype), Pc.expand(*bshape, m, m)], dim=-1)
    A_T = torch.cat([upper, lower], dim=-2)
    # the all-missing branch (:700-705) does not use B: keep its (singular) QR out of the autograd graph
    A_T_raw = A_T
    A_T = torch.where(flag[..., None, None], _eye_like(A_T).expand_as(A_T), A_T)
    B = _mT(_qr_r(_mT(A_T)))  # :679
    Fc = B[..., :p, :p]
    Fc_out = Fc
    if bool(flag.any()):  # the reported F_chol of an all-missing step is the factor of the real (singular) block
        with torch.no_grad():
            B_raw = _mT(_qr_r(_mT(A_T_raw)))
        Fc_out = torch.where(flag[..., None, None], B_raw[..., :p, :p], Fc)
    KFc = B[..., p:, :p]
    Pcf = B[..., p:, p:]
    # non-degenerate branch (:685-698); guard the triangular solves where the branch is not selected
    Fc_safe = torch.where(flag[..., None, None], _eye_like(Fc).expand_as(Fc), Fc)
    s = torch.linalg.solve_triangular(Fc_safe, v.unsqueeze(-1), upper=False)
    a_nd = a + (KFc @ s).squeeze(-1)
    inner = torch.linalg.solve_triangular(Fc_safe, s, upper=False).squeeze(-1)
    loss = (v * inner).sum(-1)
    logdet = 2.0 * torch.log(torch.abs(torch.diagonal(Fc_safe, dim1=-2, dim2=-1))).sum(-1)
    ll_nd = -0.5 * (p * (MVN_CONST + logdet) + loss)
    a_f = torch.where(flag[..., None], a.expand_as(a_nd), a_nd)
    Pcf = torch.where(flag[..., None, None], Pc.expand_as(Pcf), Pcf)
    ll = torch.where(flag, torch.zeros_like(ll_nd), ll_nd)
    Pcf = Pcf + _eye_like(Pcf) * jitter  # :536 (jitter on the factor, quirk Q4)
    # predict (:639-648)
    a_hat = _mv(T, a_f) + c
    Qc = _chol_lower(Q)
    TP, RQ = T @ Pcf, R @ Qc
    bshape = torch.broadcast_shapes(TP.shape[:-2], RQ.shape[:-2])
    M = _mT(torch.cat([TP.expand(*bshape, *TP.shape[-2:]), RQ.expand(*bshape, *RQ.shape[-2:])], dim=-1))
    Pc_hat = _mT(_qr_r(M)[..., :m, :m])
    return a_f, a_hat, y_hat, Pcf, Pc_hat, Fc_out, ll


_STEP = {"standard": standard_step, "univariate": univariate_step, "cholesky": sqrt_step}


_TV_ORDER = ("c", "d", "T", "Z", "R", "H", "Q")
_TV_NDIM = {"c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


def kalman_filter(kind, data, a0, P0, c, d, T, Z, R, H, Q, missing_fill_value=None, cov_jitter=None,
                  logp_only=False, time_varying=()):
    """Batched filter.  ``data``: (..., n, p).  Returns the 7 outputs of ``BaseFilter.build_graph`` with the
    time axis at position -2/-3 (after any batch dims), or only ``loglike_obs`` (..., n) if ``logp_only``."""
    fill = MISSING_FILL if missing_fill_value is None else missing_fill_value
    jitter = JITTER_DEFAULT if cov_jitter is None else cov_jitter
    step = _STEP[kind]
    n = data.shape[-2]
    a, P = a0, P0
    outs = [[] for _ in range(7)]
    vals = dict(zip(_TV_ORDER, (c, d, T, Z, R, H, Q)))

    def at(name, t):  # time-varying inputs carry the time axis right before their static dims
        x = vals[name]
        if name not in time_varying:
            return x
        return x.select(-_TV_NDIM[name] - 1, t)

    for t in range(n):
        res = step(data[..., t, :], a, P, *[at(k, t) for k in _TV_ORDER], fill, jitter)
        a, P = res[1], res[4]
        if logp_only:
            outs[6].append(res[6])
        else:
            for lst, val in zip(outs, res):
                lst.append(val)
    ll = torch.stack(outs[6], dim=-1)
    if logp_only:
        return ll
    filt_a, pred_a, obs_mu = (torch.stack(o, dim=-2) for o in outs[:3])
    filt_P, pred_P, obs_cov = (torch.stack(o, dim=-3) for o in outs[3:6])
    bs = pred_a.shape[:-2]
    pred_a = torch.cat([a0.expand(*bs, a0.shape[-1]).unsqueeze(-2), pred_a[..., :-1, :]], dim=-2)
    pred_P = torch.cat([P0.expand(*bs, *P0.shape[-2:]).unsqueeze(-3), pred_P[..., :-1, :, :]], dim=-3)
    if kind == "cholesky":
        sq = lambda L: L @ _mT(L)  # :733-740
        filt_P, pred_P, obs_cov = sq(filt_P), sq(pred_P), sq(obs_cov)
    return filt_a, pred_a, obs_mu, filt_P, pred_P, obs_cov, ll


NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")


def logp_and_grad(kind, data, a0, P0, c, d, T, Z, R, H, Q, ll_cotangent=None, **kw):
    """Per-series logp (...,) and gradients of sum(ll_cotangent * ll) w.r.t. the nine inputs.

    All nine inputs must already carry the full batch shape if per-series gradients are wanted (a broadcast
    input receives the gradient summed over the batch).  Inputs are numpy arrays or tensors."""
    as_t = lambda x: torch.as_tensor(x, dtype=torch.float64)
    data = as_t(data)
    params = [as_t(x).clone().requires_grad_(True) for x in (a0, P0, c, d, T, Z, R, H, Q)]
    ll = kalman_filter(kind, data, *params, logp_only=True, **kw)
    logp = ll.sum(-1)
    obj = ll.sum() if ll_cotangent is None else (ll * as_t(ll_cotangent)).sum()
    grads = torch.autograd.grad(obj, params, allow_unused=True)
    grads = [torch.zeros_like(p_) if g is None else g for g, p_ in zip(grads, params)]
    return logp.detach().numpy(), {k: g.numpy() for k, g in zip(NAMES, grads)}
