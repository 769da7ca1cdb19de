"""ctypes wrapper of oracle/kalman_c.c (compiled, OpenMP-parallel CPU oracle).  TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_c

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
_ND = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}
KINDS = {"standard": 0, "univariate": 1}
_lib = None


def lib():
    global _lib
    if _lib is None:
        # -march=native code must be built on the machine that runs it: rebuild if the cached .so is foreign
        path = build_c.build()
        _lib = C.CDLL(path)
        vp = C.c_void_p
        _lib.kalman_c_logp_grad.argtypes = ([C.c_int] * 7 + [C.c_uint] + [vp] * 11 + [C.c_double, C.c_double] + [vp] * 10
                                            + [C.c_int])
        _lib.kalman_c_logp_grad.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def logp_and_grad(kind, y, a0, P0, c, d, T, Z, R, H, Q, ll_cotangent=None, cov_jitter=1e-8, missing_fill_value=-9999.0,
                  want_grad=True):
    """Batched (leading batch axis on any input; others shared) logp [B] and gradients {name: [B, ...]}."""
    prm = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in zip(NAMES, (a0, P0, c, d, T, Z, R, H, Q))}
    y = np.ascontiguousarray(y, dtype=np.float64)
    B = None
    mask = 0
    for i, k in enumerate(NAMES):
        if prm[k].ndim == _ND[k]:
            mask |= 1 << i
        else:
            B = prm[k].shape[0]
    y_shared = int(y.ndim == 2)
    if not y_shared:
        B = y.shape[0]
    squeeze = B is None
    B = 1 if squeeze else B
    m, r = prm["R"].shape[-2:]
    p = prm["Z"].shape[-2]
    n = y.shape[-2]
    shapes = {"a0": (m,), "P0": (m, m), "c": (m,), "d": (p,), "T": (m, m), "Z": (p, m), "R": (m, r), "H": (p, p),
              "Q": (r, r)}
    logp = np.zeros(B)
    grads = {k: np.zeros((B, *shapes[k])) for k in NAMES} if want_grad else {}
    cot = None if ll_cotangent is None else np.ascontiguousarray(ll_cotangent, dtype=np.float64)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib().kalman_c_logp_grad(
        KINDS[kind], B, n, m, p, r, y_shared, mask, ptr(y), *[ptr(prm[k]) for k in NAMES], ptr(cot),
        float(cov_jitter), float(missing_fill_value), ptr(logp), *[ptr(grads.get(k)) for k in NAMES], int(want_grad))
    if rc != 0:
        raise RuntimeError(f"kalman_c_logp_grad failed ({rc})")
    if squeeze:
        return logp[0], {k: v[0] for k, v in grads.items()}
    return logp, grads
