"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against

* the golden vectors produced by the reference's own source (tests/golden/*.npz), and
* the oracle (oracle/kalman_np.py, oracle/kalman_torch.py) on seeded batched inputs.

Tolerance (BASELINE.json north_star): float64 <= 1e-9 relative on logp and on every gradient entry
(relative to max|g| of that matrix); float32 vs the float64 oracle: 1e-3 on logp, 1e-2 on gradients.
"""

import numpy as np
import pytest
from conftest import golden_names, golden_tv, load_golden
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


def _args(ins):
    return [ins["y"]] + [ins[k] for k in NAMES]


def _ill(name):
    # divide-by-jitter cases: a missing observation with p > 1 leaves F = jitter on that diagonal entry.
    # Measured on the B200 (profiles/r02_parity_margins.json, tools/parity_margins.py): 8e-15 on gradients, so these
    # cases are held to the same 1e-9 as everything else; the smoother keeps 1e-8 like its other cases.
    return ("missing" in name and any(s in name for s in ("p2", "p3", "p5")))


def _diffuse(name):
    # P0 = 1e6 I (Nile test, config C1): every implementation loses ~6 digits in the first update
    return "nile" in name or name.startswith("c1_")


def _close(actual, desired, rtol, msg):
    scale = max(np.abs(desired).max(), 1e-300)
    assert_allclose(np.asarray(actual).reshape(desired.shape), desired, rtol=0, atol=rtol * scale, err_msg=msg)


@pytest.mark.parametrize("name", golden_names())
def test_forward_outputs_match_reference_golden(eng, name):
    kind, ins, ref = load_golden(name)
    out = eng.filter(kind, *_args(ins), time_varying=golden_tv(name))
    rtol = RTOL64  # every golden case, the diffuse (P0 = 1e6 I) and the divide-by-jitter ones included (measured margins:
    # 4e-12 / 8e-15, profiles/r02_parity_margins.json)
    assert_allclose(out["loglike_obs"].sum(), ref["loglike_obs"].sum(), rtol=rtol)
    for n in OUT_NAMES:
        _close(out[n], ref[n], rtol, f"{name}:{n}")


@pytest.mark.parametrize("name", golden_names())
def test_logp_and_gradients_match_reference_golden(eng, name):
    kind, ins, ref = load_golden(name)
    logp, grads, status = eng.logp_grad(kind, *_args(ins), time_varying=golden_tv(name))
    rtol = RTOL64  # (diffuse prior: measured 2.2e-10 = F0 / H x eps, profiles/r02_parity_margins.json "nile_diffuse")
    assert_allclose(logp, ref["loglike_obs"].sum(), rtol=rtol)
    for k in NAMES:
        _close(grads[k], ref["grad_" + k], rtol, f"{name}: d/d{k}")


@pytest.mark.parametrize("name", golden_names())
def test_smoother_matches_reference_golden(eng, name):
    kind, ins, ref = load_golden(name)
    sa, sP = eng.smooth(ins["T"], ins["R"], ins["Q"], ref["filtered_states"], ref["filtered_covariances"],
                        time_varying=tuple(k for k in golden_tv(name) if k in "TRQ"))
    tol = 1e-8  # (diffuse / divide-by-jitter cases measured at 2e-12, profiles/r02_parity_margins.json "smoother_diffuse")
    lo = 0
    if _diffuse(name):
        # P0 = 1e6 I: pinv(P_hat) of a cond ~1e14 matrix decides the first few smoothed covariances; the
        # reference's own test only compares them from index 5 on (test_kalman_filter.py:243-251)
        lo = 5
        _close(sP[:5], ref["smoothed_covariances"][:5], 2e-2, f"{name}: smoothed_covariances[:5]")
    _close(sa, ref["smoothed_states"], tol, f"{name}: smoothed_states")
    _close(sP[lo:], ref["smoothed_covariances"][lo:], tol, f"{name}: smoothed_covariances")
    assert_allclose(sa[-1], ref["filtered_states"][-1])  # reference test_kalman_filter.py:216-224
    assert_allclose(sP[-1], ref["filtered_covariances"][-1])


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("shape", [(2, 1, 2), (3, 1, 1), (5, 2, 3), (15, 1, 5), (8, 3, 2)])
@pytest.mark.parametrize("missing", [0.0, 0.1])
def test_batched_logp_grad_matches_torch_oracle(eng, kind, shape, missing):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 37, 24
    cfg = synth.random_dense(B, n, m, p, r, seed=3 * m + p, missing=missing if p == 1 else 0.0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args)
    ref_logp, ref_g = oth.logp_and_grad(kind, *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"d/d{k}[{b}]")
    assert (status == 0).all()


@pytest.mark.parametrize("kind", ["standard", "univariate"])
def test_specialised_kernel_path_is_taken_for_m15(eng, kind):
    from pymc_experimental_b200 import synth

    cfg = synth.sarima(B=5, n=12)
    eng.logp_grad(kind, cfg["y"], *[cfg[k] for k in NAMES])
    assert eng.last_path == "dmma_f64"


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("B", [1, 2, 3, 7, 130])
def test_subwarp_kernels_ragged_batches_and_asymmetric_P0(eng, kind, B):
    """m = 15 register-tiled path: batch sizes that leave lanes / sub-warps idle, an asymmetric P0 (gradient
    convention: every entry independent) and missing values."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(B, 17, 15, 1, 4, seed=40 + B, missing=0.2)
    rng = np.random.default_rng(B)
    cfg["P0"] = cfg["P0"] + 0.05 * rng.standard_normal(cfg["P0"].shape)  # not symmetric any more
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter(kind, *args)
    ref = oth.kalman_filter(kind, *[__import__("torch").as_tensor(a) for a in args])
    for got, want in zip([out[n] for n in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, kind)
    logp, grads, status = eng.logp_grad(kind, *args)
    ref_logp, ref_g = oth.logp_and_grad(kind, *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"d/d{k}[{b}]")


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_thread_per_series_kernels(eng, kind, m):
    """m <= 4 register-resident path (ETS, local level): asymmetric P0, missing values, odd batch size."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    B = 203
    cfg = synth.random_dense(B, 21, m, 1, max(1, m - 1), seed=70 + m, missing=0.15)
    cfg["P0"] = cfg["P0"] + 0.05 * np.random.default_rng(m).standard_normal(cfg["P0"].shape)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter(kind, *args)
    assert eng.last_path == "thread_f64"
    ref = oth.kalman_filter(kind, *[__import__("torch").as_tensor(a) for a in args])
    for got, want in zip([out[n] for n in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, kind)
    logp, grads, status = eng.logp_grad(kind, *args)
    ref_logp, ref_g = oth.logp_and_grad(kind, *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(0, B, 17):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"d/d{k}[{b}]")


@pytest.mark.parametrize("m", [5, 6, 8, 9, 10, 12, 13, 15, 16])
def test_subwarp_kernels_all_sizes(eng, m):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    B = 11
    cfg = synth.random_dense(B, 15, m, 1, 3, seed=90 + m, missing=0.1)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    for kind in ("standard", "univariate"):
        logp, grads, status = eng.logp_grad(kind, *args)
        assert eng.last_path == ("dmma_f64" if m >= 9 else "subwarp_f64")
        ref_logp, ref_g = oth.logp_and_grad(kind, *args)
        assert_allclose(logp, ref_logp, rtol=RTOL64)
        for k in NAMES:
            for b in range(B):
                _close(grads[k][b], ref_g[k][b], RTOL64, f"{kind} m={m} d/d{k}[{b}]")


# shapes (m, p, r) instantiated for the warp-per-series DMMA square-root kernels (bkf_sqrt_warp.cu)
SQRT_WARP_SHAPES = ((8, 8, 8), (16, 8, 8), (24, 8, 8), (32, 8, 8), (16, 16, 8), (16, 16, 16), (32, 16, 16))


@pytest.mark.parametrize("shape", [(2, 1, 2), (5, 2, 3), (8, 3, 2), (12, 4, 4), (6, 6, 3), (8, 8, 8), (16, 8, 8),
                                   (24, 8, 8), (32, 8, 8), (16, 16, 8), (16, 16, 16), (24, 8, 16), (32, 16, 16)])
@pytest.mark.parametrize("missing", [0.0, 0.2])
def test_square_root_filter_batched_logp_grad(eng, shape, missing):
    """SquareRootFilter forward + QR adjoint (P0 is a factor; H, Q gradients in the lower-triangular convention)."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 9, 14
    cfg = synth.random_dense(B, n, m, p, r, seed=300 + m, missing=0.0)
    if missing:  # only whole-step missingness is defined for this filter (cholesky(W H) fails otherwise)
        rng = np.random.default_rng(m)
        cfg["y"][rng.uniform(size=(B, n)) < missing] = np.nan
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter("cholesky", *args)
    assert eng.last_path == ("sqrt_warp_dmma" if shape in SQRT_WARP_SHAPES else "sqrt_cta")
    ref = oth.kalman_filter("cholesky", *[__import__("torch").as_tensor(a) for a in args])
    for got, want in zip([out[n_] for n_ in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, "cholesky forward")
    logp, grads, status = eng.logp_grad("cholesky", *args)
    # float64 shapes made of whole 8 x 8 tiles run the warp-per-series DMMA forward sweep and the four-warp adjoint,
    # everything else the CTA kernels
    assert eng.last_path == ("sqrt_warp_dmma+adjoint_4warp" if shape in SQRT_WARP_SHAPES else "sqrt_cta")
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"cholesky d/d{k}[{b}]")  # measured 2e-14 (r02_parity_margins.json)


@pytest.mark.parametrize("case", ["dense_P0", "H_zero", "cotangent", "shared_params", "zero_Z_row"])
def test_square_root_warp_kernels_corner_cases(eng, case):
    """Branches of the warp-per-series square-root kernels the main sweep does not reach: a factor P0 that is not
    triangular (only the first step sees it), H == 0 (the reference then skips cholesky(H), kalman_filter.py:666),
    a non-trivial log-likelihood cotangent, parameters shared across the batch, and an observed series that loads on no
    state (a zero row of Z with diagonal H: that column of the update array has nothing below its diagonal, so its
    Householder reflector is the identity -- dlarfg's tau = 0 branch)."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = 16, 8, 8
    B, n = 5, 9
    cfg = synth.random_dense(B, n, m, p, r, seed=77, missing=0.0)
    cfg["y"][1, 3] = np.nan  # one skipped step
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    kw = {}
    if case == "dense_P0":
        cfg["P0"] = cfg["P0"] + 0.1 * np.random.default_rng(5).normal(size=cfg["P0"].shape)
    elif case == "H_zero":
        cfg["H"] = np.zeros_like(cfg["H"])
    elif case == "cotangent":
        kw["ll_cotangent"] = np.random.default_rng(6).normal(size=(B, n))
    elif case == "zero_Z_row":
        cfg["Z"][:, 2, :] = 0.0
        cfg["H"] = cfg["H"] * np.eye(p)
    full = {k: cfg[k] for k in NAMES}
    if case == "shared_params":  # the engine returns per-series gradients also for a parameter passed once
        for k in ("Z", "R", "c", "d"):
            cfg[k] = cfg[k][0]
            full[k] = np.ascontiguousarray(np.broadcast_to(cfg[k], (B,) + cfg[k].shape))
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("cholesky", *args, **kw)
    assert eng.last_path == "sqrt_warp_dmma+adjoint_4warp"
    ref_logp, ref_g = oth.logp_and_grad("cholesky", cfg["y"], *[full[k] for k in NAMES], **kw)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    # H == 0 with p observed series makes the filtered covariance exactly singular: its factor is jitter-sized (1e-8) in p
    # directions, and the QR pull-back (R^{-1} ... R^{-T}) amplifies rounding by that ratio in BOTH implementations
    # (measured: up to 3e-6 relative between this kernel and autograd through the oracle, 2e-5 for the CTA kernels).
    # d logp / dH is exactly zero there -- the rows of the QR input that hold H are zero, so R^T R = A^T A does not
    # move at first order -- and the oracle's value is rounding noise (1e-9 of the other gradients' scale).
    # measured (profiles/r02_parity_margins.json "square_root_corner_H_zero"): 4.7e-6 with cond(factor) = 1e9, i.e.
    # cond x eps = 2e-7 and cond^2 x eps >> 1 -- the one case that stays loose
    tol = 1e-4 if case == "H_zero" else RTOL64
    for k in NAMES:
        if case == "H_zero" and k == "H":
            assert np.all(grads[k] == 0.0)
            assert np.abs(ref_g[k]).max() < 1e-6 * max(np.abs(ref_g[j]).max() for j in NAMES)
            continue
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], tol, f"{case} d/d{k}[{b}]")


def test_ll_cotangent_is_respected(eng):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(5, 16, 4, 1, 2, seed=5)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    cot = np.random.default_rng(0).standard_normal((5, 16))
    for kind in ("standard", "univariate"):
        _, grads, _ = eng.logp_grad(kind, *args, ll_cotangent=cot)
        _, ref_g = oth.logp_and_grad(kind, *args, ll_cotangent=cot)
        for k in NAMES:
            _close(grads[k], ref_g[k], RTOL64, f"{kind} d/d{k}")


def test_shared_parameters_and_shared_y(eng):
    from oracle import kalman_np as onp
    from pymc_experimental_b200 import synth

    cfg = synth.structural(B=6, n=30)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    shared = dict(cfg, T=cfg["T"][0], Z=cfg["Z"][0], R=cfg["R"][0])
    out = eng.filter("standard", shared["y"], *[shared[k] for k in NAMES], outputs=("loglike_obs",))
    for b in range(6):
        ref = onp.kalman_filter("standard", cfg["y"], *[cfg[k][b] for k in NAMES])[-1]
        assert_allclose(out["loglike_obs"][b], ref, rtol=1e-9, atol=1e-11)


def test_device_pointers_give_the_same_answer_as_host_pointers(eng):
    import torch

    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(9, 20, 6, 1, 3, seed=9)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp_h, g_h, _ = eng.logp_grad("univariate", *args)
    targs = [torch.as_tensor(a, device="cuda") for a in args]
    logp_d, g_d, _ = eng.logp_grad("univariate", *targs)
    torch.cuda.synchronize()
    assert_allclose(logp_d.cpu().numpy(), logp_h, rtol=0, atol=0)
    for k in NAMES:
        assert_allclose(g_d[k].cpu().numpy(), g_h[k], rtol=0, atol=0)


@pytest.mark.parametrize("kind", ["standard", "univariate"])
def test_float32_within_stated_bound(eng, kind):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    cfg = synth.structural(B=8, n=60)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, _ = eng.logp_grad(kind, *args, dtype=np.float32, cov_jitter=1e-6)
    assert eng.last_path == "subwarp_f32"  # m = 15: the float32 sub-warp kernels, not the generic fallback
    ref_logp, ref_g = oth.logp_and_grad(kind, *args, cov_jitter=1e-6)
    assert_allclose(logp, ref_logp, rtol=1e-3)
    for k in NAMES:
        _close(grads[k], ref_g[k], 1e-2, f"f32 d/d{k}")


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("m", [5, 8, 9, 12, 16])
def test_float32_subwarp_kernels_every_width(eng, kind, m):
    """The float32 instantiation of the sub-warp family (bkf_subwarp.cuh with real = float; 8 lanes per series up to m = 8,
    16 above), forward outputs and logp + gradients against the float64 oracle at the stated float32 bound (1e-3 / 1e-2;
    the reference's own float32 tolerance is 1e-3, tests/statespace/test_kalman_filter.py:30-31), with missing values, an
    asymmetric P0 (first-step ASYM path) and a batch that does not fill the last warp."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    B, n = 11, 30
    cfg = synth.random_dense(B, n, m, 1, 3, seed=900 + m, missing=0.1)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args, dtype=np.float32, cov_jitter=1e-6)
    assert eng.last_path == "subwarp_f32" and logp.dtype == np.float32
    assert (status == 0).all()
    ref_logp, ref_g = oth.logp_and_grad(kind, *args, cov_jitter=1e-6)
    assert_allclose(logp, ref_logp, rtol=1e-3)
    for k in NAMES:
        _close(grads[k], ref_g[k], 1e-2, f"f32 m={m} d/d{k}")
    res = eng.filter(kind, *args, dtype=np.float32, cov_jitter=1e-6)
    ref = oth.kalman_filter(kind, *[__import__("torch").as_tensor(a) for a in args], cov_jitter=1e-6)
    for name, r in zip(("filtered_states", "predicted_states", "observed_states", "filtered_covariances",
                        "predicted_covariances", "observed_covariances", "loglike_obs"), ref):
        _close(res[name], r.numpy(), 2e-3, f"f32 m={m} {name}")


def test_filter_smooth_fused_call(eng):
    from oracle import kalman_np as onp
    from pymc_experimental_b200 import synth

    cfg = synth.structural(B=4, n=40)
    ll, sa, sP, status = eng.filter_smooth("standard", cfg["y"], *[cfg[k] for k in NAMES])
    for b in range(4):
        outs = onp.kalman_filter("standard", cfg["y"], *[cfg[k][b] for k in NAMES])
        ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], outs[0], outs[3])
        assert_allclose(ll[b], outs[-1], rtol=1e-9, atol=1e-11)
        _close(sa[b], ra, 1e-8, "smoothed_states")
        _close(sP[b], rP, 1e-8, "smoothed_covariances")


@pytest.mark.parametrize("m", [5, 8, 9, 11, 13, 15, 16])
def test_subwarp_smoother_matches_oracle(eng, m):
    from oracle import kalman_np as onp
    from pymc_experimental_b200 import synth

    B, n = 7, 19
    cfg = synth.random_dense(B, n, m, 1, 3, seed=200 + m, missing=0.1)
    ll, sa, sP, status = eng.filter_smooth("standard", cfg["y"], *[cfg[k] for k in NAMES])
    assert eng.last_path == ("dmma_f64" if m >= 9 else "subwarp_f64")
    for b in range(B):
        outs = onp.kalman_filter("standard", cfg["y"][b], *[cfg[k][b] for k in NAMES])
        ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], outs[0], outs[3])
        assert_allclose(ll[b], outs[-1], rtol=1e-9, atol=1e-11)
        _close(sa[b], ra, 1e-8, f"m={m} smoothed_states[{b}]")
        _close(sP[b], rP, 1e-8, f"m={m} smoothed_covariances[{b}]")


def test_ill_conditioned_smoother_falls_back_to_eigen_pinv(eng):
    """P_hat with pivots spanning > 1e13: the Cholesky smoother flags the series and exactly that series is redone with
    the eigen-decomposition pinv (the reference's SVD pinv semantics); the other series of the batch keep the bits of the
    fast kernel, and no host round trip is involved (the flagged indices stay on the device)."""
    from oracle import kalman_np as onp

    from pymc_experimental_b200.engine import STATUS_SMOOTHER_ILL

    m, n, B = 6, 5, 9
    rng = np.random.default_rng(3)
    T = np.eye(m)
    R = np.eye(m)[:, :2]
    Q = np.eye(2) * 1e-3
    fa = rng.standard_normal((B, n, m))
    good = np.diag([2.0, 1.0, 1.0, 0.5, 1.0, 1.0])
    fP = np.broadcast_to(good, (B, n, m, m)).copy()
    ill = (2, 7)
    for b in ill:
        fP[b] = np.diag([1e9, 1.0, 1.0, 1e-6, 1.0, 1.0])
    sa, sP, status = eng.smooth(T, R, Q, fa, fP, return_status=True)
    assert [int(x) & STATUS_SMOOTHER_ILL for x in status] == [STATUS_SMOOTHER_ILL if b in ill else 0 for b in range(B)]
    for b in range(B):
        ra, rP = onp.kalman_smoother(T, R, Q, fa[b], fP[b])
        tol = 1e-5 if b in ill else 1e-8
        _close(sa[b], ra, tol, f"smoothed_states[{b}]")
        _close(sP[b], rP, tol, f"smoothed_covariances[{b}]")
    keep = [b for b in range(B) if b not in ill]
    sa2, sP2 = eng.smooth(T, R, Q, fa[keep], fP[keep])
    assert np.array_equal(sa[keep], sa2) and np.array_equal(sP[keep], sP2)
    sa1, sP1 = eng.smooth(T, R, Q, fa[ill[0]], fP[ill[0]])  # the single-series call of the reference's own usage
    assert np.array_equal(sa1, sa[ill[0]]) and np.array_equal(sP1, sP[ill[0]])


def test_empty_batch_and_single_step(eng):
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(3, 1, 4, 1, 2, seed=1)
    logp, grads, _ = eng.logp_grad("standard", cfg["y"], *[cfg[k] for k in NAMES])
    assert logp.shape == (3,) and np.isfinite(logp).all()
    empty = {k: cfg[k][:0] for k in NAMES}
    logp0, _, _ = eng.logp_grad("standard", cfg["y"][:0], *[empty[k] for k in NAMES])
    assert logp0.shape == (0,)


def test_unsupported_configurations_fail_loudly(eng):
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import BkfError

    cfg = synth.random_dense(2, 5, 3, 1, 2, seed=2)
    bad = dict(cfg, T=np.broadcast_to(cfg["T"][:, None], (2, 5, 3, 3)))  # time-varying T
    with pytest.raises(BkfError):
        eng.filter("standard", bad["y"], *[bad[k] for k in NAMES])


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("tv", [("Z",), ("T", "Q"), ("c", "d", "T", "Z", "R", "H", "Q")])
@pytest.mark.parametrize("shape", [(3, 1, 2), (5, 2, 3), (15, 1, 5)])
def test_time_varying_matrices_batched(eng, kind, tv, shape):
    """Inputs with a time axis (filters/utilities.py:8-47): batched logp, gradients
    (which keep the time axis) and filter + smoother against the oracles.  Shared (unbatched) time-varying inputs
    are covered by the golden cases."""
    import torch

    from oracle import kalman_np as onp
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 9, 14
    cfg = synth.random_dense_tv(B, n, m, p, r, tv, seed=11 * m + p + len(tv), missing=0.1 if p == 1 else 0.0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args, time_varying=tv)
    # a time axis on c, d, Z, H only stays on the fast kernels (p = 1): thread per series (m <= 4), sub-warp (5..8), DMMA (9..16)
    obs_side = p == 1 and set(tv) <= {"c", "d", "Z", "H"}
    want = "generic_warp"
    if obs_side:
        want = "thread_f64" if m <= 4 else ("subwarp_f64" if m <= 8 else ("dmma_f64" if m <= 16 else "generic_warp"))
    assert eng.last_path == want
    ref_logp, ref_g = oth.logp_and_grad(kind, *args, time_varying=tv)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        assert grads[k].shape == cfg[k].shape
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"{kind} tv={tv} d/d{k}[{b}]")
    out = eng.filter(kind, *args, time_varying=tv)
    ref = oth.kalman_filter(kind, *[torch.as_tensor(a) for a in args], time_varying=tv)
    for got, want in zip([out[k] for k in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, f"{kind} tv={tv}")
    tv_s = tuple(k for k in tv if k in "TRQ")
    sa, sP = eng.smooth(cfg["T"], cfg["R"], cfg["Q"], out["filtered_states"], out["filtered_covariances"],
                        time_varying=tv_s)
    for b in range(0, B, 4):
        ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], out["filtered_states"][b],
                                     out["filtered_covariances"][b])
        _close(sa[b], ra, 1e-8, f"smoothed_states[{b}]")
        _close(sP[b], rP, 1e-8, f"smoothed_covariances[{b}]")


@pytest.mark.parametrize("kind", ["standard", "univariate"])
@pytest.mark.parametrize("tv", [("Z",), ("c", "d"), ("H",), ("c", "d", "Z", "H")])
@pytest.mark.parametrize("m", [2, 3, 4, 5, 7, 8, 9, 15, 16])
@pytest.mark.parametrize("shared", [False, True])
def test_time_varying_observation_side_on_the_fast_kernels(eng, kind, tv, m, shared):
    """Row f-4 on the fast path: a time axis on c, d, Z and / or H (the RegressionComponent pattern Z[t],
    structural.py:1601-1617; time-varying intercepts and measurement noise) is re-read per step by the DMMA kernels
    and by the thread-per-series kernels (m <= 4: local level / trend + regression) themselves -- logp, gradients (which
    keep the time axis, also for an input shared by the batch), the seven filter outputs and filter + smoother against the
    oracles, with missing observations."""
    import torch

    from oracle import kalman_np as onp
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    B, n, r = 7, 21, min(3, m)
    fast_path = "thread_f64" if m <= 4 else ("subwarp_f64" if m <= 8 else "dmma_f64")
    cfg = synth.random_dense_tv(B, n, m, 1, r, tv, seed=5 * m + len(tv), missing=0.15)
    ref_cfg = dict(cfg)
    if shared:  # the time-varying inputs once for the whole batch: [n, ...] instead of [B, n, ...]
        for k in tv:
            cfg[k] = np.ascontiguousarray(cfg[k][0])
            ref_cfg[k] = np.repeat(cfg[k][None], B, 0)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    ref_args = [ref_cfg["y"]] + [ref_cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad(kind, *args, time_varying=tv)
    assert eng.last_path == fast_path and (status == 0).all()
    ref_logp, ref_g = oth.logp_and_grad(kind, *ref_args, time_varying=tv)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        assert grads[k].shape == ref_cfg[k].shape, k  # per-series gradients, time axis kept
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"{kind} tv={tv} d/d{k}[{b}]")
    out = eng.filter(kind, *args, time_varying=tv)
    assert eng.last_path == fast_path
    ref = oth.kalman_filter(kind, *[torch.as_tensor(a) for a in ref_args], time_varying=tv)
    for got, want in zip([out[k] for k in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, f"{kind} tv={tv}")
    if kind == "standard":
        ll, sa, sP, _ = eng.filter_smooth(kind, *args, time_varying=tv)
        if m > 4:
            assert eng.last_path == fast_path
        for b in (0, B - 1):
            ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], out["filtered_states"][b],
                                         out["filtered_covariances"][b])
            _close(sa[b], ra, 1e-8, f"smoothed_states[{b}]")
            _close(sP[b], rP, 1e-8, f"smoothed_covariances[{b}]")


def test_time_varying_observation_side_float32_thread_kernels(eng):
    """The float32 instantiation of the same path (m <= 4): a regression design row Z[t] and d[t], stated float32 bound."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    tv = ("Z", "d")
    cfg = synth.random_dense_tv(6, 25, 3, 1, 2, tv, seed=3, missing=0.1)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("standard", *args, time_varying=tv, dtype=np.float32, cov_jitter=1e-6)
    assert eng.last_path == "thread_f32" and logp.dtype == np.float32
    ref_logp, ref_g = oth.logp_and_grad("standard", *args, time_varying=tv, cov_jitter=1e-6)
    assert_allclose(logp, ref_logp, rtol=1e-3)
    for k in NAMES:
        assert grads[k].shape == cfg[k].shape
        _close(grads[k], ref_g[k], 1e-2, f"f32 tv d/d{k}")


@pytest.mark.parametrize("shape", [(5, 2, 3), (16, 8, 8)])
@pytest.mark.parametrize("tv", [("Z",), ("c", "d"), ("T", "H"), ("R",), ("Q",), ("R", "Q"),
                                ("c", "d", "T", "Z", "R", "H", "Q")])
def test_time_varying_square_root_filter(eng, tv, shape):
    """SquareRootFilter with a time axis on any of c, d, T, Z, R, H, Q (filters/utilities.py:8-47): filter outputs, logp
    and gradients (per-step records for the time-varying inputs, the Cholesky pull-backs of H_t and Q_t per step, the
    other of R / Q accumulated) against autograd through the oracle; one whole-step missing observation."""
    import torch

    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 5, 9
    cfg = synth.random_dense_tv(B, n, m, p, r, tv, seed=13 * m + len(tv))
    cfg["y"][1, 4] = np.nan
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter("cholesky", *args, time_varying=tv)
    assert eng.last_path == "sqrt_cta"
    ref = oth.kalman_filter("cholesky", *[torch.as_tensor(a) for a in args], time_varying=tv)
    for got, want in zip([out[k] for k in OUT_NAMES], ref):
        _close(got, want.numpy(), RTOL64, f"cholesky tv={tv}")
    logp, grads, status = eng.logp_grad("cholesky", *args, time_varying=tv)
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args, time_varying=tv)
    assert_allclose(logp, ref_logp, rtol=RTOL64)
    for k in NAMES:
        assert grads[k].shape == cfg[k].shape
        for b in range(B):
            _close(grads[k][b], ref_g[k][b], RTOL64, f"cholesky tv={tv} d/d{k}[{b}]")


@pytest.mark.parametrize("device", [False, True])
def test_batch_is_sliced_when_the_workspace_would_not_fit(eng, device):
    """KalmanEngine.logp_grad asks the library how much device workspace a call needs
    (bkf_logp_grad_workspace_bytes) and processes the batch in slices when that exceeds its budget (70 % of the
    device by default; forced here).  Series are independent: the sliced result is bit-identical."""
    import torch

    from pymc_experimental_b200 import synth

    B, n = 37, 11
    cfg = synth.sarima(B=B, n=n)
    cot = np.random.default_rng(3).standard_normal((B, n))
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    if device:
        args = [torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") for a in args]
        cot = torch.as_tensor(cot, device="cuda:0")
    to_np = lambda x: x.cpu().numpy() if device else x
    logp1, g1, st1 = eng.logp_grad("univariate", *args, ll_cotangent=cot)
    launches = eng.launch_count
    eng.max_workspace_bytes = 10 * n * 292 * 8  # room for ~10 series of the fragment-order trajectory
    try:
        logp2, g2, st2 = eng.logp_grad("univariate", *args, ll_cotangent=cot)
    finally:
        eng.max_workspace_bytes = None
    assert eng.launch_count - launches >= 8  # at least four slices, forward + backward each
    np.testing.assert_array_equal(to_np(logp1), to_np(logp2))
    np.testing.assert_array_equal(to_np(st1), to_np(st2))
    for k in NAMES:
        np.testing.assert_array_equal(to_np(g1[k]), to_np(g2[k]), err_msg=k)
    free, total = eng.device_memory()
    assert 0 < free <= total


@pytest.mark.parametrize("device", [False, True])
def test_compact_parameterisation_equals_the_dense_call(eng, device):
    """draws that differ in a few matrix entries ship only those entries
    (bkf_filter_logp_grad_compact).  The expanded matrices are the dense call's inputs, so logp is bit-identical and
    g_u is the dense gradient gathered at the mapped positions."""
    import torch

    from pymc_experimental_b200 import synth

    B, n = 29, 40
    cfg = synth.sarima(B=B, n=n)
    # SARIMA(1,1,1)x(1,0,1,12): phi, Phi, -Phi*phi in T; theta, Theta, Theta*theta in R; sigma^2 in Q (SARIMAX.py:383-536)
    positions = [("T", (1, 1)), ("T", (12, 1)), ("T", (13, 1)), ("R", (2, 0)), ("R", (13, 0)), ("R", (14, 0)),
                 ("Q", (0, 0))]
    base = {k: np.array(cfg[k][0]) for k in NAMES}
    u = np.stack([cfg[name][(slice(None), *idx)] for name, idx in positions], axis=1)
    cot = np.random.default_rng(9).standard_normal((B, n))
    dense = [cfg["y"]] + [cfg[k] for k in NAMES]
    y, uu, cc, bb = cfg["y"], u, cot, base
    if device:
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
        dense, y, uu, cc, bb = [t(a) for a in dense], t(y), t(u), t(cot), {k: t(v) for k, v in base.items()}
    to_np = lambda x: x.cpu().numpy() if device else x
    logp_d, g_d, st_d = eng.logp_grad("univariate", *dense, ll_cotangent=cc)
    logp_c, g_u, st_c = eng.logp_grad_compact("univariate", y, bb, positions, uu, ll_cotangent=cc)
    assert eng.last_path == "dmma_f64"
    np.testing.assert_array_equal(to_np(logp_c), to_np(logp_d))
    np.testing.assert_array_equal(to_np(st_c), to_np(st_d))
    for k, (name, idx) in enumerate(positions):
        np.testing.assert_array_equal(to_np(g_u)[:, k], to_np(g_d[name])[(slice(None), *idx)], err_msg=f"{name}{idx}")
    from pymc_experimental_b200.engine import BkfError

    with pytest.raises(BkfError, match="unique"):
        eng.logp_grad_compact("univariate", y, bb, positions + [("T", (1, 1))],
                              np.concatenate([u, u[:, :1]], axis=1) if not device else torch.cat([uu, uu[:, :1]], 1))


@pytest.mark.parametrize("device", [False, True])
def test_theta_program_equals_the_compact_call_chained_by_hand(eng, device):
    """The rest of row f-1 (bkf_filter_logp_grad_theta): the free parameters go in, the entries are evaluated and the chain
    rule applied on the device.  SARIMA(1,1,1)x(1,0,1,12) (models/SARIMAX.py:404-458): logp equals the compact call fed with
    u(theta) from the NumPy restatement of the program, g_theta = J^T g_u."""
    import torch

    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth, theta_maps

    B, n = 23, 60
    cfg = synth.sarima(B=1, n=n)
    base = {k: np.array(cfg[k][0]) for k in NAMES}
    positions, terms, ntheta = theta_maps.sarima_111_101_12()
    rng = np.random.default_rng(77)
    theta = np.column_stack([rng.uniform(-0.8, 0.8, B), rng.uniform(-0.8, 0.8, B), rng.uniform(-0.5, 0.5, B),
                             rng.uniform(-0.5, 0.5, B), np.exp(rng.normal(0.0, 0.3, B))])
    u = theta_maps.evaluate(terms, theta, len(positions))
    J = theta_maps.jacobian(terms, theta, len(positions))
    cot = rng.standard_normal((B, n))
    wrap = (lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")) if device else (lambda a: a)
    to_np = lambda x: x.cpu().numpy() if device else x
    y, bb = wrap(cfg["y"]), {k: wrap(v) for k, v in base.items()}
    logp_c, g_u, st_c = eng.logp_grad_compact("univariate", y, bb, positions, wrap(u), ll_cotangent=wrap(cot))
    logp_t, g_t, st_t = eng.logp_grad_theta("univariate", y, bb, positions, terms, wrap(theta), ll_cotangent=wrap(cot))
    np.testing.assert_allclose(to_np(logp_t), to_np(logp_c), rtol=1e-13)
    np.testing.assert_array_equal(to_np(st_t), to_np(st_c))
    ref = np.einsum("bk,bkt->bt", to_np(g_u), J)
    _close(to_np(g_t), ref, 1e-12, "g_theta")
    # and against the dense call + the oracle's gradient for one draw
    dense = {k: np.repeat(base[k][None], B, 0) for k in NAMES}
    for k, (name, idx) in enumerate(positions):
        dense[name][(slice(None), *idx)] = u[:, k]
    ref_logp, _ = oth.logp_and_grad("univariate", cfg["y"], *[dense[k][:3] for k in NAMES])
    logp_plain, _, _ = eng.logp_grad_theta("univariate", y, bb, positions, terms, wrap(theta), want_grad=False)
    np.testing.assert_allclose(to_np(logp_plain)[:3], ref_logp, rtol=RTOL64)


def test_theta_program_with_every_function_code(eng):
    """cos / sin / exp / 1 - x / x^2 and three-factor terms: a stochastic cycle (models/structural.py:1220-1237) inside the
    C3 structural model, plus an artificial entry that uses the remaining codes; central differences of the device's own
    logp tie g_theta to the function it is the gradient of."""
    from pymc_experimental_b200 import synth, theta_maps
    from pymc_experimental_b200.engine import FN_EXP, FN_ID, FN_ONE_MINUS

    B, n = 9, 50
    cfg = synth.structural(B=1, n=n)
    base = {k: np.array(cfg[k][0]) for k in NAMES}
    positions, terms, ntheta = theta_maps.structural_cycle(first_state=13, shock=3)
    positions = positions + [("H", (0, 0)), ("c", (0,))]
    terms = terms + [(6, 0.5, [(3, FN_EXP), (4, FN_ONE_MINUS), (3, FN_EXP)]), (7, 0.1, [(4, FN_ID)]), (7, 0.05, [])]
    ntheta = 5
    rng = np.random.default_rng(5)
    theta = np.column_stack([rng.uniform(0.7, 0.98, B), rng.uniform(0.1, 0.5, B), rng.normal(-1.0, 0.2, B),
                             rng.normal(-1.0, 0.2, B), rng.uniform(0.1, 0.6, B)])
    u = theta_maps.evaluate(terms, theta, len(positions))
    J = theta_maps.jacobian(terms, theta, len(positions))
    logp_c, g_u, _ = eng.logp_grad_compact("standard", cfg["y"], base, positions, u)
    logp_t, g_t, st = eng.logp_grad_theta("standard", cfg["y"], base, positions, terms, theta)
    assert (st == 0).all()
    np.testing.assert_allclose(logp_t, logp_c, rtol=1e-12)
    _close(g_t, np.einsum("bk,bkt->bt", g_u, J), 1e-11, "g_theta")
    h = 1e-6
    for j in range(ntheta):
        e = np.zeros(ntheta)
        e[j] = h
        lp, _, _ = eng.logp_grad_theta("standard", cfg["y"], base, positions, terms, theta + e, want_grad=False)
        lm, _, _ = eng.logp_grad_theta("standard", cfg["y"], base, positions, terms, theta - e, want_grad=False)
        fd = (lp - lm) / (2 * h)
        np.testing.assert_allclose(g_t[:, j], fd, rtol=2e-5, atol=2e-5 * np.abs(g_t).max())


@pytest.mark.parametrize("shape", [(3, 1), (6, 3), (15, 1), (16, 4), (32, 16)])
@pytest.mark.parametrize("device", [False, True])
def test_stationary_initialisation_matches_the_oracle(eng, shape, device):
    """a0 = (I - T)^-1 c and P0 = solve_discrete_lyapunov(T, R Q R^T)
    (models/SARIMAX.py:369-381, VARMAX.py:379-393) for a batch, and their pull-back, against oracle/stationary.py."""
    import torch

    from oracle import stationary as ost

    m, r = shape
    B = 7
    rng = np.random.default_rng(50 + m)
    T = rng.standard_normal((B, m, m))
    rho = np.array([0.5, 0.9, 0.97, 0.2, 0.8, 0.99, 0.6])
    T *= (rho / np.abs(np.linalg.eigvals(T)).max(axis=1))[:, None, None]
    R, c = rng.standard_normal((B, m, r)), rng.standard_normal((B, m))
    L = rng.standard_normal((B, r, r))
    Q = L @ np.swapaxes(L, -1, -2) + 0.1 * np.eye(r)
    ab, Pb = rng.standard_normal((B, m)), rng.standard_normal((B, m, m))
    wrap = (lambda a: torch.as_tensor(a, device="cuda:0")) if device else (lambda a: a)
    to_np = lambda x: x.cpu().numpy() if device else x
    a0, P0, status = eng.stationary_init(wrap(c), wrap(T), wrap(R), wrap(Q))
    assert eng.last_path == "stationary_warp"
    assert (to_np(status) == 0).all()
    g = eng.stationary_init_grad(wrap(T), wrap(R), wrap(Q), a0, P0, wrap(ab), wrap(Pb))
    for b in range(B):
        ra, rP, rg = ost.stationary_init_grad(c[b], T[b], R[b], Q[b], ab[b], Pb[b])
        _close(to_np(a0)[b], ra, 1e-9, f"a0[{b}]")
        _close(to_np(P0)[b], rP, 1e-9, f"P0[{b}]")
        for k in ("c", "T", "R", "Q"):
            _close(to_np(g[k])[b], rg[k], 1e-8, f"d/d{k}[{b}]")
    # a unit root is not stationary: flagged, not silently wrong
    Tu = np.eye(m)[None]
    _, _, st = eng.stationary_init(c[:1], Tu, R[:1], Q[:1])
    assert st[0] == 8


@pytest.mark.parametrize("m", [3, 8, 15, 16, 32])
def test_mvn_sample_matches_oracle_and_reproduces_covariance(eng, m):
    """Row f-3 (SequenceMvNormal / MvNormalSVD draw): the kernel's x = mu + L z equals the oracle's Cholesky-with-dropped-
    pivots restatement, L L^T reproduces cov (recovered by feeding unit vectors as z), for full-rank and rank-deficient
    covariances, host and device pointers."""
    import torch

    from oracle import sampling as osm

    rng = np.random.default_rng(100 + m)
    B, n = 5, 7
    A = rng.normal(size=(B, n, m, m))
    cov = A @ A.transpose(0, 1, 3, 2) + 0.05 * np.eye(m)
    rk = max(1, m // 3)
    cov[1] = A[1][..., :rk] @ A[1][..., :rk].transpose(0, 2, 1)  # one series with rank-deficient covariances
    mu = rng.normal(size=(B, n, m))
    z = rng.standard_normal((B, n, m))
    x, status = eng.sample_mvn(mu, cov, z)
    assert eng.last_path == "mvn_chol_subwarp"
    assert int(np.abs(status).sum()) == 0
    ref = osm.sample(mu, cov, z)
    scale = np.abs(ref).max()
    assert_allclose(x, ref, rtol=0, atol=1e-10 * scale)
    xd, _ = eng.sample_mvn(*[torch.as_tensor(a, device="cuda") for a in (mu, cov, z)])
    assert_allclose(xd.cpu().numpy(), np.asarray(x), rtol=0, atol=0)
    # columns of L from unit variates: L L^T == cov
    L = np.empty((B, n, m, m))
    for k in range(m):
        e = np.zeros((B, n, m))
        e[..., k] = 1.0
        L[..., k] = np.asarray(eng.sample_mvn(np.zeros_like(mu), cov, e)[0])
    assert_allclose(L @ L.transpose(0, 1, 3, 2), cov, rtol=0, atol=1e-11 * np.abs(cov).max())
    # an indefinite matrix is flagged, not silently accepted
    bad = cov.copy()
    bad[2, 3] = -np.eye(m)
    _, st = eng.sample_mvn(mu, bad, z)
    assert st[2] != 0 and st[0] == 0


def test_sample_sequence_mvnormal_host_mirror(eng):
    """sampling.sample_sequence_mvnormal: draws live where the inputs do, shape (..., T, m), reproducible from the rng."""
    from pymc_experimental_b200.sampling import sample_sequence_mvnormal

    rng = np.random.default_rng(4)
    A = rng.normal(size=(2, 3, 6, 4, 4))
    covs = A @ A.transpose(0, 1, 2, 4, 3) + 0.1 * np.eye(4)
    mus = rng.normal(size=(2, 3, 6, 4))
    x1 = sample_sequence_mvnormal(mus, covs, rng=np.random.default_rng(9), engine=eng)
    x2 = sample_sequence_mvnormal(mus, covs, rng=np.random.default_rng(9), engine=eng)
    assert x1.shape == mus.shape and np.array_equal(x1, x2)
    z = np.random.default_rng(9).standard_normal(mus.shape)
    L = np.linalg.cholesky(covs)
    assert_allclose(x1, mus + np.einsum("...ij,...j->...i", L, z), rtol=0, atol=1e-11 * np.abs(x1).max())


@pytest.mark.parametrize("shape", [(3, 1, 1), (15, 1, 5), (8, 3, 2), (32, 16, 16)])
def test_simulate_matches_oracle(eng, shape):
    """Row f-3: LinearGaussianStateSpace.rv_op (unconditional simulation) against its restatement, batched and shared
    parameters, a rank-deficient Q, host and device pointers."""
    import torch

    from oracle import sampling as osm
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 6, 11
    cfg = synth.random_dense(B, n, m, p, r, seed=40 + m)
    cfg["Q"][2] = cfg["Q"][2][:, :1] @ cfg["Q"][2][:, :1].T  # rank one
    rng = np.random.default_rng(m)
    zs = [rng.standard_normal(sh) for sh in ((B, m), (B, p), (B, n, r), (B, n, p))]
    xs, ys, status = eng.simulate(*[cfg[k] for k in NAMES], *zs)
    assert eng.last_path == "simulate_warp" and int(np.abs(status).sum()) == 0
    for b in range(B):
        rx, ry = osm.simulate(*[cfg[k][b] for k in NAMES], *[z[b] for z in zs])
        assert_allclose(xs[b], rx, rtol=0, atol=1e-10 * np.abs(rx).max())
        assert_allclose(ys[b], ry, rtol=0, atol=1e-10 * np.abs(ry).max())
    shared = dict(cfg, T=cfg["T"][0], Z=cfg["Z"][0], R=cfg["R"][0], Q=cfg["Q"][0])
    xs2, ys2, _ = eng.simulate(*[torch.as_tensor(shared[k], device="cuda") for k in NAMES],
                               *[torch.as_tensor(z, device="cuda") for z in zs])
    rx, ry = osm.simulate(*[(shared[k] if shared[k].ndim < cfg[k].ndim else shared[k][1]) for k in NAMES],
                          *[z[1] for z in zs])
    assert_allclose(xs2[1].cpu().numpy(), rx, rtol=0, atol=1e-10 * np.abs(rx).max())
    assert_allclose(ys2[1].cpu().numpy(), ry, rtol=0, atol=1e-10 * np.abs(ry).max())


@pytest.mark.parametrize("tv", [("Z",), ("c", "d", "T", "Z", "R", "H", "Q")])
def test_simulate_with_time_varying_matrices(eng, tv):
    """sequence_names of LinearGaussianStateSpace.rv_op: step t uses x[t], the initial observation Z[0] and H[0]."""
    from oracle import sampling as osm
    from pymc_experimental_b200 import synth

    m, p, r = 6, 2, 3
    B, n = 4, 8
    cfg = synth.random_dense_tv(B, n, m, p, r, tv, seed=71)
    rng = np.random.default_rng(8)
    zs = [rng.standard_normal(sh) for sh in ((B, m), (B, p), (B, n, r), (B, n, p))]
    xs, ys, status = eng.simulate(*[cfg[k] for k in NAMES], *zs, time_varying=tv)
    assert int(np.abs(status).sum()) == 0
    for b in range(B):
        rx, ry = osm.simulate(*[cfg[k][b] for k in NAMES], *[z[b] for z in zs], time_varying=tv)
        assert_allclose(xs[b], rx, rtol=0, atol=1e-10 * np.abs(rx).max())
        assert_allclose(ys[b], ry, rtol=0, atol=1e-10 * np.abs(ry).max())


def test_simulate_then_filter_is_calibrated(eng):
    """Size-independent property tying the simulator to the filter: for data simulated from the model, the standardised
    one-step innovations of the Kalman filter are white with unit variance, so sum_t v_t^2 / F_t ~ chi2(T) per series."""
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.sampling import simulate_statespace

    cfg = synth.structural(B=2048, n=5)
    steps = 200
    out = simulate_statespace(*[cfg[k] for k in NAMES], steps=steps, rng=np.random.default_rng(2), engine=eng,
                              append_x0=False)
    assert out.shape == (2048, steps, 16)
    y = out[..., 15:]
    # y_1.. are generated from x_0 ~ N(a0, P0) one transition earlier: the matching prior for y_1 is the prediction of x_1
    a1 = np.einsum("bij,bj->bi", cfg["T"], cfg["a0"]) + cfg["c"]
    RQR = cfg["R"] @ cfg["Q"] @ cfg["R"].transpose(0, 2, 1)
    P1 = cfg["T"] @ cfg["P0"] @ cfg["T"].transpose(0, 2, 1) + RQR
    f = eng.filter("standard", y, a1, P1, *[cfg[k] for k in NAMES[2:]],
                   outputs=("observed_states", "observed_covariances"))
    v = y - f["observed_states"]
    q = (v[..., 0] ** 2 / f["observed_covariances"][..., 0, 0]).sum(-1)  # chi2(steps): mean steps, var 2 steps
    assert abs(q.mean() - steps) < 5 * np.sqrt(2 * steps / 2048)
    assert abs(q.var() / (2 * steps) - 1.0) < 0.2


@pytest.mark.parametrize("kind,shape", [("standard", (15, 1, 5)), ("univariate", (6, 3, 2)), ("cholesky", (16, 8, 8))])
def test_fused_filter_smooth_sample_equals_the_three_calls(eng, kind, shape):
    """bkf_filter_smooth_sample (sample_conditional_posterior in one call) is bit-identical to filter_smooth followed by
    sample_mvn on its outputs; only the draws and the log-likelihood come back."""
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 7, 12
    cfg = synth.random_dense(B, n, m, p, r, seed=5 + m, missing=0.1 if p == 1 else 0.0)
    if kind == "cholesky":
        cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    z = np.random.default_rng(1).standard_normal((B, n, m))
    x, ll, status = eng.sample_conditional_posterior(kind, *args, z)
    ll2, sa, sP, _ = eng.filter_smooth(kind, *args)
    x2, _ = eng.sample_mvn(sa, sP, z)
    np.testing.assert_array_equal(np.asarray(x), np.asarray(x2))
    np.testing.assert_array_equal(np.asarray(ll), np.asarray(ll2))
    assert int(np.abs(status).sum()) == 0
    from pymc_experimental_b200.sampling import sample_conditional_posterior

    x3, ll3 = sample_conditional_posterior(kind, *args, rng=np.random.default_rng(1), engine=eng)
    np.testing.assert_array_equal(np.asarray(x3), np.asarray(x))  # same generator state -> same variates -> same draws


def test_sampling_entry_points_in_float32(eng):
    """The float32 instantiations of bkf_mvn_sample / bkf_simulate against the float64 oracle (single-precision bound)."""
    from oracle import sampling as osm
    from pymc_experimental_b200 import synth

    rng = np.random.default_rng(21)
    B, n, m, p, r = 4, 6, 8, 2, 3
    A = rng.normal(size=(B, n, m, m))
    cov = A @ A.transpose(0, 1, 3, 2) + 0.5 * np.eye(m)
    mu = rng.normal(size=(B, n, m))
    z = rng.standard_normal((B, n, m))
    x, status = eng.sample_mvn(mu, cov, z, dtype=np.float32)
    assert x.dtype == np.float32 and int(np.abs(status).sum()) == 0
    ref = osm.sample(mu, cov, z)
    assert_allclose(x, ref, rtol=0, atol=2e-5 * np.abs(ref).max())
    cfg = synth.random_dense(B, n, m, p, r, seed=3)
    zs = [rng.standard_normal(sh) for sh in ((B, m), (B, p), (B, n, r), (B, n, p))]
    xs, ys, status = eng.simulate(*[cfg[k] for k in NAMES], *zs, dtype=np.float32)
    assert xs.dtype == np.float32 and int(np.abs(status).sum()) == 0
    for b in range(B):
        rx, ry = osm.simulate(*[cfg[k][b] for k in NAMES], *[zz[b] for zz in zs])
        assert_allclose(xs[b], rx, rtol=0, atol=1e-4 * np.abs(rx).max())
        assert_allclose(ys[b], ry, rtol=0, atol=1e-4 * np.abs(ry).max())


def test_in_kernel_generator_equals_the_oracle_variates(eng):
    """Row f-3 with a NULL variate pointer: the kernels draw their standard normals themselves (Philox4x32-10 +
    Box-Muller keyed by seed / offset).  Feeding the oracle's restatement of that generator through the explicit-z
    path gives the same draws (log / sincospi differ from NumPy's by rounding only), for every batch split."""
    from oracle import philox
    from pymc_experimental_b200 import synth

    B, n, m = 37, 11, 15
    cfg = synth.structural(B=B, n=n)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    ll, sa, sP, _ = eng.filter_smooth("standard", *args)
    seed, off = 0x1234_5678_9ABC_DEF0, 5
    x, st = eng.sample_mvn(sa, sP, seed=seed, offset=off)
    z = philox.standard_normal(seed, off, 0, B * n * m).reshape(B, n, m)
    xz, _ = eng.sample_mvn(sa, sP, z)
    assert_allclose(x, xz, rtol=0, atol=1e-12 * np.abs(xz).max())
    # variate i does not depend on the batch the call covers: a slice of the batch with the matching offset
    b0 = 8  # (b0 * n * m is even: a counter yields two variates)
    xs, _ = eng.sample_mvn(sa[b0:20], sP[b0:20], seed=seed, offset=off + b0 * n * m // 2)
    assert np.array_equal(xs, x[b0:20])
    # fused filter + smoother + draw
    xf, llf, _ = eng.sample_conditional_posterior("standard", *args, seed=seed, offset=off)
    assert np.array_equal(xf, x) and np.array_equal(llf, ll)
    # unconditional simulation: the four variate arrays are streams 1..4 of the same (seed, offset)
    p, r = cfg["Z"].shape[-2], cfg["R"].shape[-1]
    prm = [cfg[k] for k in NAMES]
    xs1, ys1, _ = eng.simulate(*prm, batch=B, steps=n, seed=seed, offset=off)
    zs = [philox.standard_normal(seed, off, s_, cnt).reshape(shape) for s_, cnt, shape in
          ((1, B * m, (B, m)), (2, B * p, (B, p)), (3, B * n * r, (B, n, r)), (4, B * n * p, (B, n, p)))]
    xs2, ys2, _ = eng.simulate(*prm, *zs)
    assert_allclose(xs1, xs2, rtol=0, atol=1e-12 * np.abs(xs2).max())
    assert_allclose(ys1, ys2, rtol=0, atol=1e-12 * np.abs(ys2).max())
    with pytest.raises(Exception):
        eng.sample_mvn(sa, sP)  # neither variates nor a seed


def test_comm_entry_points_single_rank(eng):
    """bkf_comm_* (the NCCL all-gather of [logp | grads] through the C ABI, SURVEY section 8 rows b / e) with a communicator
    of one rank: the gather of a packed logp + gradient buffer returns that buffer.  The two-rank case is
    tests/test_comm_two_gpus.py (needs two GPUs)."""
    import torch

    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

    cfg = synth.sarima(B=16, n=30)
    dev = torch.device("cuda", 0)
    dargs = [torch.as_tensor(cfg["y"], device=dev)] + [torch.as_tensor(cfg[k], device=dev) for k in NAMES]
    layout, total = eng.packed_layout(16, 15, 1, 1)
    packed = torch.empty(total, dtype=torch.float64, device=dev)
    eng.logp_grad("univariate", *dargs, packed_out=packed)
    eng.comm_init(1, 0, KalmanEngine.comm_unique_id())
    out = eng.comm_allgather(packed)
    torch.cuda.synchronize()
    assert out.shape == (1, total)
    assert torch.equal(out[0], packed)
    eng.comm_destroy()


def test_overlapped_host_call_is_bit_identical(eng):
    """KalmanEngine.logp_grad_overlapped: the batch in slices on two handles / two streams / two host threads so that the
    copies hide under the kernels; series are independent, so every output equals the single call bit for bit (shared
    parameters, a cotangent and a batch that does not divide evenly included)."""
    from pymc_experimental_b200 import synth

    B, n = 333, 40
    cfg = synth.sarima(B=B, n=n)
    cfg["Z"] = cfg["Z"][0]  # a shared parameter
    cot = np.random.default_rng(4).standard_normal((B, n))
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("univariate", *args, ll_cotangent=cot)
    for chunks in (2, 3, 5):
        lo, go, so = eng.logp_grad_overlapped("univariate", *args, chunks=chunks, ll_cotangent=cot)
        np.testing.assert_array_equal(lo, logp)
        np.testing.assert_array_equal(so, status)
        for k in NAMES:
            np.testing.assert_array_equal(go[k], grads[k], err_msg=f"chunks={chunks} {k}")
