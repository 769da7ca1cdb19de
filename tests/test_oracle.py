"""CPU tests pinning the oracle (oracle/kalman_np.py, oracle/kalman_torch.py).

1. against golden vectors produced by executing the reference's own source (tests/golden/generate_golden.py);
2. against the reference's scipy-MVN identity test (tests/statespace/test_distributions.py:96-119);
3. against a textbook, jitter-free Kalman filter + RTS smoother standing in for statsmodels, at the
   reference's own tolerance atol=rtol=1e-6 (tests/statespace/test_kalman_filter.py:30-31,232-269);
4. the reference's invariants (test_kalman_filter.py:216-224, 277-301).
"""

import numpy as np
import pytest
from conftest import golden_names, golden_tv, load_golden
from numpy.testing import assert_allclose
from scipy import stats

from oracle import kalman_np as onp
from oracle import kalman_torch as oth

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
OUT_NAMES = ("filtered_states", "predicted_states", "observed_states", "filtered_covariances",
             "predicted_covariances", "observed_covariances", "loglike_obs")
ALL = golden_names()


def _args(ins):
    return [ins["y"]] + [ins[k] for k in NAMES]


def _ill_conditioned(name):
    # partially-missing p>1 cases divide by the 1e-8 jitter: values ~1e6..1e8
    return "missing" in name and ("p2" in name or "p3" in name or "p5" in name)


@pytest.mark.parametrize("name", ALL)
def test_numpy_oracle_matches_reference_outputs(name):
    kind, ins, ref = load_golden(name)
    outs = onp.kalman_filter(kind, *_args(ins))
    rtol = 1e-6 if _ill_conditioned(name) else 1e-10
    for n, o in zip(OUT_NAMES, outs):
        scale = max(1.0, np.abs(ref[n]).max())
        assert_allclose(o.reshape(ref[n].shape), ref[n], rtol=rtol, atol=rtol * scale, err_msg=f"{name}:{n}")
    sm_a, sm_P = onp.kalman_smoother(ins["T"], ins["R"], ins["Q"], outs[0], outs[3])
    stol = 1e-5 if ("nile" in name or _ill_conditioned(name)) else 1e-8  # pinv of cond~1e14 matrices on Nile
    assert_allclose(sm_a, ref["smoothed_states"], rtol=stol, atol=stol * max(1, np.abs(sm_a).max()))
    assert_allclose(sm_P, ref["smoothed_covariances"], rtol=stol, atol=stol * max(1, np.abs(sm_P).max()))


@pytest.mark.parametrize("name", ALL)
def test_torch_oracle_matches_reference_logp_and_grad(name):
    kind, ins, ref = load_golden(name)
    logp, grads = oth.logp_and_grad(kind, *_args(ins), time_varying=golden_tv(name))
    rtol = 1e-6 if _ill_conditioned(name) else 1e-9
    assert_allclose(logp, ref["loglike_obs"].sum(), rtol=rtol)
    for k in NAMES:
        g_ref = ref["grad_" + k]
        scale = max(np.abs(g_ref).max(), 1e-300)
        assert_allclose(grads[k], g_ref, rtol=0, atol=rtol * scale, err_msg=f"{name}: d/d{k}")


def test_torch_oracle_batched_equals_per_series():
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(4, 12, 4, 2, 3, seed=7)
    for kind in ("standard", "univariate", "cholesky"):
        logp, grads = oth.logp_and_grad(kind, cfg["y"], *[cfg[k] for k in NAMES])
        for b in range(4):
            lp1, g1 = oth.logp_and_grad(kind, cfg["y"][b], *[cfg[k][b] for k in NAMES])
            assert_allclose(logp[b], lp1, rtol=1e-12)
            for k in NAMES:
                assert_allclose(grads[k][b], g1[k], rtol=1e-9, atol=1e-12)


def test_gradient_matches_finite_differences():
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(1, 10, 3, 2, 2, seed=11)
    args = [cfg["y"][0]] + [cfg[k][0] for k in NAMES]
    for kind in ("standard", "univariate", "cholesky"):
        _, grads = oth.logp_and_grad(kind, *args)
        for k_idx, k in enumerate(NAMES):
            x = args[1 + k_idx]
            it = np.nditer(x, flags=["multi_index"])
            for _ in it:
                idx = it.multi_index
                h = 1e-6
                xp, xm = x.copy(), x.copy()
                xp[idx] += h
                xm[idx] -= h
                g = grads[k][idx]
                if k in ("P0", "H", "Q") and idx[0] != idx[1] and kind != "cholesky":
                    # LAPACK posv reads one triangle: only symmetric perturbations are well defined on the
                    # NumPy side; compare with the sum of the two entry-wise gradients
                    xp[idx[::-1]] += h
                    xm[idx[::-1]] -= h
                    g = g + grads[k][idx[::-1]]
                ap, am = list(args), list(args)
                ap[1 + k_idx], am[1 + k_idx] = xp, xm
                fd = (onp.kalman_logp(kind, *ap) - onp.kalman_logp(kind, *am)) / (2 * h)
                assert_allclose(g, fd, rtol=2e-5, atol=2e-6, err_msg=f"{kind} d/d{k}{idx}")


@pytest.mark.parametrize("kind", ["standard", "univariate", "cholesky"])
def test_loglike_vectors_agree_with_scipy_mvn(kind):
    """reference: tests/statespace/test_distributions.py:96-119 (k_endog = 1, atol=rtol=1e-5 there)."""
    _, ins, _ = load_golden("nile_standard_miss0")
    if kind == "cholesky":
        ins = dict(ins, P0=np.linalg.cholesky(ins["P0"]))
    outs = onp.kalman_filter(kind, *_args(ins))
    obs_mu, obs_cov, ll = outs[2], outs[5], outs[6]
    for t in range(len(ll)):
        ref = stats.multivariate_normal.logpdf(ins["y"][t], mean=obs_mu[t], cov=obs_cov[t])
        assert_allclose(ll[t], ref, rtol=1e-7, atol=1e-7)


def _textbook_filter_smoother(y, a0, P0, T, Z, R, H, Q):
    """Durbin & Koopman (2012) eqs 4.24 / 4.44: plain covariance filter + RTS smoother, no jitter, missing
    observations skipped.  Independent of the oracle; stands in for statsmodels' KalmanFilter."""
    n, m = len(y), len(a0)
    a, P = a0.copy(), P0.copy()
    fa, fP, pa, pP, ll = [], [], [], [], []
    RQR = R @ Q @ R.T
    for t in range(n):
        pa.append(a), pP.append(P)
        if np.isnan(y[t]).all():
            af, Pf = a, P
            ll.append(0.0)
        else:
            v = y[t] - Z @ a
            F = Z @ P @ Z.T + H
            Fi = np.linalg.inv(F)
            K = P @ Z.T @ Fi
            af = a + K @ v
            Pf = P - K @ Z @ P
            ll.append(-0.5 * (len(v) * np.log(2 * np.pi) + np.log(np.linalg.det(F)) + v @ Fi @ v))
        fa.append(af), fP.append(Pf)
        a, P = T @ af, T @ Pf @ T.T + RQR
    fa, fP, pa, pP = map(np.array, (fa, fP, pa, pP))
    sa, sP = fa.copy(), fP.copy()
    for t in range(n - 2, -1, -1):
        Pp = T @ fP[t] @ T.T + RQR
        G = fP[t] @ T.T @ np.linalg.inv(Pp)
        sa[t] = fa[t] + G @ (sa[t + 1] - T @ fa[t])
        sP[t] = fP[t] + G @ (sP[t + 1] - Pp) @ G.T
    return fa, pa, sa, fP, pP, sP, np.array(ll)


@pytest.mark.parametrize("kind", ["standard", "univariate", "cholesky"])
@pytest.mark.parametrize("n_missing", [0, 5])
def test_filters_match_textbook_filter_on_nile(kind, n_missing):
    """reference: tests/statespace/test_kalman_filter.py:232-269 (statsmodels replaced by a textbook filter;
    same tolerances, smoothed covariances compared from index 5 on as there)."""
    _, ins, _ = load_golden(f"nile_standard_miss{n_missing}")
    fa, pa, sa, fP, pP, sP, ll = _textbook_filter_smoother(
        ins["y"], ins["a0"], ins["P0"], ins["T"], ins["Z"], ins["R"], ins["H"], ins["Q"])
    if kind == "cholesky":
        ins = dict(ins, P0=np.linalg.cholesky(ins["P0"]))
    outs = onp.kalman_filter(kind, *_args(ins))
    sm_a, sm_P = onp.kalman_smoother(ins["T"], ins["R"], ins["Q"], outs[0], outs[3])
    tol = dict(atol=1e-6, rtol=1e-6)
    assert_allclose(outs[0], fa, **tol)
    assert_allclose(outs[1], pa, **tol)
    assert_allclose(outs[3], fP, **tol)
    assert_allclose(outs[4], pP, **tol)
    assert_allclose(outs[6], ll, **tol)
    assert_allclose(outs[6].sum(), ll.sum(), **tol)
    assert_allclose(sm_a, sa, **tol)
    assert_allclose(sm_P[5:], sP[5:], **tol)


@pytest.mark.parametrize("kind", ["standard", "univariate", "cholesky"])
def test_last_smoother_is_last_filtered_and_covariances_psd(kind):
    """reference: test_kalman_filter.py:216-224 and :277-301."""
    _, ins, _ = load_golden("nile_standard_miss5")
    if kind == "cholesky":
        ins = dict(ins, P0=np.linalg.cholesky(ins["P0"]))
    outs = onp.kalman_filter(kind, *_args(ins))
    sm_a, sm_P = onp.kalman_smoother(ins["T"], ins["R"], ins["Q"], outs[0], outs[3])
    assert_allclose(sm_a[-1], outs[0][-1])
    assert_allclose(sm_P[-1], outs[3][-1])
    if kind == "univariate":  # the reference's PSD test skips the univariate filter (test_kalman_filter.py:274)
        return
    for cov in (outs[3], outs[4], sm_P):  # filtered, predicted, smoothed covariances
        assert (np.linalg.eigvalsh(0.5 * (cov + np.swapaxes(cov, -1, -2))) > 0).all()
        assert_allclose(cov, np.swapaxes(cov, -1, -2), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", [n for n in ALL if "cholesky" not in n and not n.startswith("tv_")])
def test_c_oracle_matches_reference_logp_and_grad(name):
    """oracle/kalman_c.c (the compiled CPU baseline) against the vectors produced from the reference's source."""
    from oracle import kalman_c

    kind, ins, ref = load_golden(name)
    logp, grads = kalman_c.logp_and_grad(kind, *_args(ins))
    rtol = 1e-6 if _ill_conditioned(name) else (1e-7 if ("nile" in name or name.startswith("c1_")) else 1e-9)
    assert_allclose(logp, ref["loglike_obs"].sum(), rtol=rtol)
    for k in NAMES:
        g_ref = ref["grad_" + k]
        scale = max(np.abs(g_ref).max(), 1e-300)
        assert_allclose(grads[k], g_ref, rtol=0, atol=rtol * scale, err_msg=f"{name}: d/d{k}")


def test_c_oracle_batched_and_threaded():
    from oracle import kalman_c
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense(13, 12, 5, 2, 3, seed=17, missing=0.0)
    cot = np.random.default_rng(1).standard_normal((13, 12))
    for kind in ("standard", "univariate"):
        logp, grads = kalman_c.logp_and_grad(kind, cfg["y"], *[cfg[k] for k in NAMES], ll_cotangent=cot)
        ref_logp, ref_g = oth.logp_and_grad(kind, cfg["y"], *[cfg[k] for k in NAMES], ll_cotangent=cot)
        assert_allclose(logp, ref_logp, rtol=1e-11)
        for k in NAMES:
            assert_allclose(grads[k], ref_g[k], rtol=0, atol=1e-9 * np.abs(ref_g[k]).max())


@pytest.mark.parametrize("tv", [("Z",), ("c", "d", "T", "Z", "R", "H", "Q")])
def test_univariate_time_varying_oracles_agree(tv):
    """The reference's UnivariateFilter cannot run with time-varying inputs (kalman_filter.py:791 skips
    unpack_args), so this combination is pinned on the two independent restatements agreeing with each other and,
    for p = 1, with the StandardFilter (same density up to jitter placement)."""
    from pymc_experimental_b200 import synth

    cfg = synth.random_dense_tv(1, 16, 4, 1, 2, tv, seed=5)
    args = [cfg["y"][0]] + [cfg[k][0] for k in NAMES]
    ll_np = onp.kalman_filter("univariate", *args)[-1]
    ll_th = oth.kalman_filter("univariate", *[__import__("torch").as_tensor(a) for a in args], logp_only=True,
                              time_varying=tv).numpy()
    assert_allclose(ll_np, ll_th, rtol=1e-11)
    ll_std = onp.kalman_filter("standard", *args)[-1]
    assert_allclose(ll_np, ll_std, rtol=1e-6)


def test_stationary_init_oracle_direct_equals_bilinear_and_gradient_matches_finite_differences():
    """oracle/stationary.py: the two solve_discrete_lyapunov methods the reference selects between agree, the result
    satisfies P0 = T P0 T^T + R Q R^T, and the torch gradient matches central differences."""
    from oracle import stationary as ost

    rng = np.random.default_rng(21)
    m, r = 6, 3
    T = rng.standard_normal((m, m))
    T *= 0.9 / np.abs(np.linalg.eigvals(T)).max()
    R, c = rng.standard_normal((m, r)), rng.standard_normal(m)
    L = rng.standard_normal((r, r))
    Q = L @ L.T + 0.1 * np.eye(r)
    a0, P0 = ost.stationary_init_np(c, T, R, Q, "direct")
    _, P0b = ost.stationary_init_np(c, T, R, Q, "bilinear")
    assert_allclose(P0, P0b, rtol=1e-10, atol=1e-12)
    assert_allclose(P0, T @ P0 @ T.T + R @ Q @ R.T, rtol=1e-10, atol=1e-12)
    assert_allclose(a0, T @ a0 + c, rtol=1e-10, atol=1e-12)
    ab, Pb = rng.standard_normal(m), rng.standard_normal((m, m))
    _, _, g = ost.stationary_init_grad(c, T, R, Q, ab, Pb)
    def f(c, T, R, Q):
        a, P = ost.stationary_init_np(c, T, R, Q)
        return (a * ab).sum() + (P * Pb).sum()

    args = {"c": c, "T": T, "R": R, "Q": Q}
    for k, x in args.items():
        for idx in [tuple(rng.integers(0, s) for s in x.shape) for _ in range(4)]:
            xp, xm = x.copy(), x.copy()
            xp[idx] += 1e-6
            xm[idx] -= 1e-6
            fd = (f(**{**args, k: xp}) - f(**{**args, k: xm})) / 2e-6
            assert_allclose(g[k][idx], fd, rtol=1e-6, atol=1e-7)


def test_sampling_oracle_factors_reproduce_the_covariance():
    """Row f-3: the reference's SVD square root (numpy multivariate_normal, method="svd") and the Cholesky-with-dropped-
    pivots square root of the CUDA kernel both reproduce cov -- full rank and rank deficient -- so the two draws have the
    same distribution; and a Monte-Carlo check of the oracle's draw against numpy's own sampler."""
    from oracle import sampling as osm

    rng = np.random.default_rng(12)
    m = 7
    A = rng.normal(size=(m, m))
    full = A @ A.T + 0.1 * np.eye(m)
    low = A[:, :3] @ A[:, :3].T  # rank 3
    for cov in (full, low):
        S = osm.svd_factor(cov)
        L = osm.chol_psd(cov)
        np.testing.assert_allclose(S.T @ S, cov, rtol=0, atol=1e-12 * np.abs(cov).max())
        np.testing.assert_allclose(L @ L.T, cov, rtol=0, atol=1e-12 * np.abs(cov).max())
        assert np.all(np.triu(L, 1) == 0)
    assert (np.abs(np.diag(osm.chol_psd(low))) > 0).sum() == 3
    # the oracle's draws and numpy's SVD-method draws agree in mean and covariance (50 000 draws, 4 sigma bounds)
    n = 50_000
    mu = rng.normal(size=m)
    z = rng.standard_normal((n, m))
    ours = mu + z @ osm.chol_psd(full).T
    ref = np.random.default_rng(3).multivariate_normal(mu, full, size=n, method="svd")
    se = np.sqrt(np.diag(full) / n)
    assert np.all(np.abs(ours.mean(0) - ref.mean(0)) < 6 * se)
    c1, c2 = np.cov(ours.T), np.cov(ref.T)
    assert np.abs(c1 - c2).max() < 0.05 * np.abs(full).max()


def test_simulation_oracle_follows_the_reference_recursion():
    """oracle/sampling.simulate against an independent plain-NumPy transcription of LinearGaussianStateSpace.rv_op's
    step_fn (distributions.py:239-263): Cholesky factors of PD matrices, no d in the first observation mean, and the
    time-varying selection (step t uses x[t], the initial observation Z[0], H[0])."""
    from oracle import sampling as osm

    rng = np.random.default_rng(31)
    m, p, r, n = 5, 2, 3, 7

    def spd(k):
        A = rng.normal(size=(k, k))
        return A @ A.T + 0.3 * np.eye(k)

    a0, P0 = rng.normal(size=m), spd(m)
    c, d = rng.normal(size=(n, m)), rng.normal(size=(n, p))
    T, Z, R = 0.4 * rng.normal(size=(n, m, m)), rng.normal(size=(n, p, m)), rng.normal(size=(n, m, r))
    H, Q = np.stack([spd(p) for _ in range(n)]), np.stack([spd(r) for _ in range(n)])
    z0, zy0 = rng.standard_normal(m), rng.standard_normal(p)
    zs, zo = rng.standard_normal((n, r)), rng.standard_normal((n, p))
    for tv in ((), ("c", "d", "T", "Z", "R", "H", "Q")):
        sel = lambda name, x, t: x[t] if name in tv else x[0]
        args = [a0, P0] + [x if name in tv else x[0] for name, x in zip("cdTZRHQ", (c, d, T, Z, R, H, Q))]
        xs, ys = osm.simulate(*args, z0, zy0, zs, zo, time_varying=tv)
        chol = np.linalg.cholesky
        x = a0 + chol(P0) @ z0
        y = sel("Z", Z, 0) @ x + chol(sel("H", H, 0)) @ zy0
        np.testing.assert_allclose(xs[0], x, rtol=1e-12)
        np.testing.assert_allclose(ys[0], y, rtol=1e-12)
        for t in range(n):
            x = sel("c", c, t) + sel("T", T, t) @ x + sel("R", R, t) @ (chol(sel("Q", Q, t)) @ zs[t])
            y = sel("d", d, t) + sel("Z", Z, t) @ x + chol(sel("H", H, t)) @ zo[t]
            np.testing.assert_allclose(xs[t + 1], x, rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(ys[t + 1], y, rtol=1e-11, atol=1e-12)


def test_philox_oracle_known_answers_and_moments():
    """The counter-based generator of the sampling kernels (oracle/philox.py): Philox4x32-10 against the known-answer
    vectors of Random123 (kat_vectors, philox4x32 10 rounds), and the Box-Muller variates are standard normal."""
    from oracle import philox

    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for counter, key, want in kat:
        got = philox.philox4x32_10(np.array(counter, dtype=np.uint64)[None], key)[0]
        assert tuple(int(v) for v in got) == want
    z = philox.standard_normal(seed=20261020, offset=7, stream=0, count=400_000)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3.0) < 5e-2
    # a variate is a function of (seed, offset, stream, i) only: offsets shift the sequence by two variates per count
    z2 = philox.standard_normal(seed=20261020, offset=8, stream=0, count=100)
    assert np.array_equal(z2, z[2:102])
    assert not np.array_equal(philox.standard_normal(20261020, 7, 1, 100), z[:100])
