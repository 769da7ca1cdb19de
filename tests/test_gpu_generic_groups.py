"""GPU parity of the generic Standard / Univariate kernels in their CTA-per-series builds (bkf_generic.cu, W = 2, 4, 8 warps
per series; W = 1 is the one-warp mapping), against the torch / numpy oracles.  These kernels serve what has no
specialised kernel: several observed series (the reference's default filter for BayesianVARMAX, models/VARMAX.py) and one
observed series with more than 16 states (kalman_filter.py:594-617, 769-815).  BKF_GEN_W pins the width the launcher would
otherwise choose from the arena size, so every build is exercised at every shape.
"""

import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
OUT_NAMES = ("filtered_states", "predicted_states", "observed_states", "filtered_covariances",
             "predicted_covariances", "observed_covariances", "loglike_obs")
RTOL64 = 1e-9


@pytest.fixture(scope="module")
def eng():
    from pymc_experimental_b200.engine import KalmanEngine

    return KalmanEngine(0)


@pytest.fixture
def width(request):
    old = os.environ.get("BKF_GEN_W")
    os.environ["BKF_GEN_W"] = str(request.param)
    yield request.param
    if old is None:
        os.environ.pop("BKF_GEN_W", None)
    else:
        os.environ["BKF_GEN_W"] = old


def _close(actual, desired, rtol, msg):
    scale = max(np.abs(desired).max(), 1e-300)
    assert_allclose(np.asarray(actual).reshape(desired.shape), desired, rtol=0, atol=rtol * scale, err_msg=msg)


@pytest.mark.parametrize("width", [1, 2, 4, 8], indirect=True)
@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("shape", [(16, 8, 8), (20, 3, 4), (27, 1, 3), (32, 16, 16), (40, 2, 5)])
def test_group_kernels_logp_grad_and_outputs(eng, width, kind, shape):
    import torch

    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 5, 11
    cfg = synth.random_dense(B, n, m, p, r, seed=17 * m + p, missing=0.15 if p == 1 else 0.0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args)
    assert eng.last_path in ("generic_warp", "generic_cta") and (status == 0).all()
    ref_logp, ref_g = oth.logp_and_grad(kind, *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"W={width} {kind} {shape} d/d{k}[{b}]")
    out = eng.filter(kind, *args)
    ref = oth.kalman_filter(kind, *[torch.as_tensor(a) for a in args])
    for name, got, want in zip(OUT_NAMES, [out[k] for k in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, f"W={width} {kind} {shape} {name}")


@pytest.mark.parametrize("width", [2, 8], indirect=True)
@pytest.mark.parametrize("kind", ["standard", "univariate"])
def test_group_kernels_time_varying_and_cotangent(eng, width, kind):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = 18, 2, 3
    B, n = 4, 9
    tv = ("c", "d", "T", "Z", "R", "H", "Q")
    cfg = synth.random_dense_tv(B, n, m, p, r, tv, seed=5)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    cot = np.random.default_rng(3).normal(size=(B, n))
    logp, grads, status = eng.logp_grad(kind, *args, time_varying=tv, ll_cotangent=cot)
    assert eng.last_path in ("generic_warp", "generic_cta")
    ref_logp, ref_g = oth.logp_and_grad(kind, *args, time_varying=tv, ll_cotangent=cot)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        assert grads[k].shape == cfg[k].shape
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"W={width} {kind} tv d/d{k}[{b}]")


@pytest.mark.parametrize("kind", ["standard", "univariate"])
def test_group_kernels_all_widths_agree_and_repeat_bit_for_bit(eng, kind):
    """The builds differ only in who computes which element (and in the three CTA-wide sums of the univariate adjoint):
    their results agree to rounding, and each build repeats itself bit for bit (no atomics anywhere)."""
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(7, 13, 24, 4, 6, seed=2)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    res = {}
    old = os.environ.get("BKF_GEN_W")
    try:
        for w in (1, 2, 4, 8):
            os.environ["BKF_GEN_W"] = str(w)
            logp, grads, _ = eng.logp_grad(kind, *args)
            logp2, grads2, _ = eng.logp_grad(kind, *args)
            assert np.array_equal(logp, logp2) and all(np.array_equal(grads[k], grads2[k]) for k in NAMES)
            res[w] = (logp, grads)
    finally:
        if old is None:
            os.environ.pop("BKF_GEN_W", None)
        else:
            os.environ["BKF_GEN_W"] = old
    for w in (2, 4, 8):
        assert_allclose(res[w][0], res[1][0], rtol=1e-12)
        for k in NAMES:
            _close(res[w][1][k], res[1][1][k], 1e-11, f"W={w} {kind} d/d{k}")


@pytest.mark.parametrize("m,r", [(20, 20), (27, 27), (36, 8)])
def test_generic_smoother_cholesky_solve_and_pinv_paths_match_the_oracle(eng, m, r):
    """State widths without a specialised smoother (m > 16): the generic smoother solves P_hat W = T P by Cholesky when the
    pivots allow it and keeps the eigen-decomposition pinv otherwise (kalman_smoother.py:107-126, `pt.linalg.pinv`); both
    against the numpy oracle (numpy's SVD pinv) on systems with a well-conditioned P_hat."""
    from oracle import kalman_np as onp
    from pymc_experimental_b200 import synth

    B, n = 4, 12
    cfg = synth.random_dense(B, n, m, 2, r, seed=100 + m, missing=0.0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter("standard", *args, outputs=("filtered_states", "filtered_covariances"))
    old, old_w = os.environ.get("BKF_GEN_SMOOTH_PINV"), os.environ.get("BKF_GEN_W")
    try:
        for mode, width in (("chol", None), ("pinv", None), ("chol", 1), ("chol", 4), ("pinv", 8)):
            if mode == "pinv":
                os.environ["BKF_GEN_SMOOTH_PINV"] = "1"
            else:
                os.environ.pop("BKF_GEN_SMOOTH_PINV", None)
            if width is None:  # the launcher's own choice of warps per series
                os.environ.pop("BKF_GEN_W", None)
            else:
                os.environ["BKF_GEN_W"] = str(width)
            sa, sP, status = eng.smooth(cfg["T"], cfg["R"], cfg["Q"], out["filtered_states"], out["filtered_covariances"],
                                        return_status=True)
            assert (status == 0).all()
            for b in range(B):
                ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], out["filtered_states"][b],
                                             out["filtered_covariances"][b])
                _close(sa[b], ra, 1e-8, f"{mode} W={width} smoothed_states[{b}]")
                _close(sP[b], rP, 1e-8, f"{mode} W={width} smoothed_covariances[{b}]")
    finally:
        for key, val in (("BKF_GEN_SMOOTH_PINV", old), ("BKF_GEN_W", old_w)):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("shape", [(6, 6, 3), (8, 8, 4), (16, 16, 4), (5, 7, 2)])
def test_as_many_observed_series_as_states(eng, kind, shape):
    """p about m (or above it): the transposed scratch of the update's solves no longer fits X..Y and the kernels take the
    in-place substitution and a separate solve for F^-1 (bkf_generic.cu solve_scratch_fits); both builds (m < 16 and m >= 16)."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 5, 10
    cfg = synth.random_dense(B, n, m, p, r, seed=7 * m + p, missing=0.0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args)
    assert eng.last_path in ("generic_warp", "generic_cta") and (status == 0).all()
    ref_logp, ref_g = oth.logp_and_grad(kind, *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"{kind} {shape} d/d{k}[{b}]")
