"""GPU tests at the FULL sizes of the BASELINE.json configs, through size-independent properties of the domain.

The oracle cannot run 8 million state-steps in a test, so each full-size run is checked by

* a spot check: a handful of series picked from the full batch must equal the oracle run on just those series
  (catches any batch-indexing / scheduling defect that only shows with every SM busy and several waves of CTAs);
* the directional derivative: for a random direction delta in parameter space,
  (logp(theta + h delta) - logp(theta - h delta)) / 2h  ==  <grad(theta), delta>  for EVERY series of the batch, with
  the two shifted log-likelihoods evaluated by the CUDA forward kernels themselves;
* filter identities: for one observed series the Univariate and Standard filters are the same density up to jitter
  placement; the last smoothed moment equals the last filtered one (reference: test_kalman_filter.py:216-224) and
  smoothed covariances are symmetric with a positive diagonal (:277-301).

Inputs live on the device (torch tensors): these tests also exercise the device-pointer side of the C ABI.
"""

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")


@pytest.fixture(scope="module")
def eng():
    from pymc_experimental_b200.engine import KalmanEngine

    return KalmanEngine(0)


def _dev(cfg):
    import torch

    return {k: torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64, device="cuda:0") for k, v in cfg.items()}


def _spot_check(kind, cfg, logp, grads, picks, rtol=1e-9):
    from oracle import kalman_torch as oth

    y = cfg["y"] if cfg["y"].ndim == 2 else cfg["y"][picks]
    ref_logp, ref_g = oth.logp_and_grad(kind, y, *[cfg[k][picks] for k in NAMES])
    assert_allclose(logp[picks], ref_logp, rtol=rtol)
    for k in NAMES:
        for i, b in enumerate(picks):
            scale = max(np.abs(ref_g[k][i]).max(), 1e-300)
            assert_allclose(grads[k][b], ref_g[k][i], rtol=0, atol=rtol * scale, err_msg=f"d/d{k}[{b}]")


def _directional_derivative(eng, kind, dcfg, grads, names, h, rtol, seed=0):
    """<grad, delta> against a central difference of the engine's own logp, for every series."""
    import torch

    gen = torch.Generator(device="cuda:0").manual_seed(seed)
    delta = {k: torch.randn(dcfg[k].shape, generator=gen, dtype=torch.float64, device="cuda:0") for k in names}
    for k in ("H", "Q", "P0"):  # keep the covariances symmetric along the path
        if k in delta:
            delta[k] = 0.5 * (delta[k] + delta[k].transpose(-1, -2))
    analytic = sum((grads[k] * delta[k]).reshape(delta[k].shape[0], -1).sum(-1) for k in names)

    def logp_at(sign):
        args = {k: (dcfg[k] + sign * h * delta[k]) if k in delta else dcfg[k] for k in NAMES}
        lp, _, _ = eng.logp_grad(kind, dcfg["y"], *[args[k] for k in NAMES], grads=())
        return lp

    numeric = (logp_at(+1.0) - logp_at(-1.0)) / (2.0 * h)
    err = (analytic - numeric).abs()
    scale = analytic.abs().clamp_min(1.0)
    assert float((err / scale).max()) < rtol, float((err / scale).max())


def test_c2_sarima_univariate_full_size(eng):
    """configs[1]: SARIMA(1,1,1)x(1,0,1,12), 16 384 draws x T=500, UnivariateFilter logp + grad."""
    from pymc_experimental_b200 import synth

    cfg = synth.sarima(B=16384, n=500)
    dcfg = _dev(cfg)
    logp, grads, status = eng.logp_grad("univariate", dcfg["y"], *[dcfg[k] for k in NAMES])
    assert eng.last_path == "dmma_f64"
    assert int(status.abs().sum()) == 0
    picks = np.array([0, 1, 4097, 8191, 12345, 16383])
    _spot_check("univariate", cfg, logp.cpu().numpy(), {k: v.cpu().numpy() for k, v in grads.items()}, picks)
    _directional_derivative(eng, "univariate", dcfg, grads, ("T", "R", "Q", "Z", "c", "d", "a0"), h=1e-6, rtol=2e-5)
    # same density as the StandardFilter for one observed series; the two place the 1e-8 jitter differently (F + eps
    # vs P - K K^T F + eps I), which moves logp by up to ~4e-6 relative on the draws where H = 0 makes F small
    logp_std, _, _ = eng.logp_grad("standard", dcfg["y"], *[dcfg[k] for k in NAMES], grads=())
    assert_allclose(logp.cpu().numpy(), logp_std.cpu().numpy(), rtol=2e-5)


def test_c3_structural_filter_smoother_full_size(eng):
    """configs[2] per GPU: structural trend + seasonal(12) + cycle, 8 192 draws x T=240, filter + smoother."""
    from oracle import kalman_np as onp
    from pymc_experimental_b200 import synth

    cfg = synth.structural(B=8192, n=240)
    dcfg = _dev(cfg)
    ll, sa, sP, status = eng.filter_smooth("standard", dcfg["y"], *[dcfg[k] for k in NAMES])
    assert eng.last_path == "dmma_f64"
    assert int((status & 3).abs().sum()) == 0
    for b in (0, 4095, 8191):
        outs = onp.kalman_filter("standard", cfg["y"], *[cfg[k][b] for k in NAMES])
        ra, rP = onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], outs[0], outs[3])
        assert_allclose(ll[b].cpu().numpy(), outs[-1], rtol=1e-9, atol=1e-11)
        assert_allclose(sa[b].cpu().numpy(), ra, rtol=0, atol=1e-8 * np.abs(ra).max())
        assert_allclose(sP[b].cpu().numpy(), rP, rtol=0, atol=1e-8 * np.abs(rP).max())
    # identities over the whole batch
    out = eng.filter("standard", dcfg["y"], *[dcfg[k] for k in NAMES],
                     outputs=("filtered_states", "filtered_covariances"))
    assert float((sa[:, -1] - out["filtered_states"][:, -1]).abs().max()) == 0.0
    assert float((sP[:, -1] - out["filtered_covariances"][:, -1]).abs().max()) == 0.0
    asym = (sP - sP.transpose(-1, -2)).abs().amax()
    assert float(asym) <= 1e-12 * float(sP.abs().amax())
    assert float(sP.diagonal(dim1=-2, dim2=-1).amin()) > 0.0


def test_c5_ets_standard_full_size(eng):
    """configs[4]: ETS(A,Ad,N), 1 000 000 independent series x T=100, StandardFilter logp + grad."""
    from pymc_experimental_b200 import synth

    cfg = synth.ets(B=1_000_000, n=100)
    dcfg = _dev(cfg)
    logp, grads, status = eng.logp_grad("standard", dcfg["y"], *[dcfg[k] for k in NAMES])
    assert eng.last_path == "thread_f64"
    assert int(status.abs().sum()) == 0
    picks = np.array([0, 31, 32, 500_000, 999_999])
    _spot_check("standard", cfg, logp.cpu().numpy(), {k: v.cpu().numpy() for k, v in grads.items()}, picks)
    # the truncation error of the central difference is O(h^2) times the third derivative, which is large along T for
    # the near-unit-root draws (measured: the error falls 100x per 10x in h, down to 8e-6 at h = 1e-7; rounding
    # enters at ~1e-7)
    _directional_derivative(eng, "standard", dcfg, grads, ("T", "R", "Q", "H", "Z", "c", "d", "a0", "P0"), h=1e-7,
                            rtol=1e-4)


def test_c4_varmax_square_root_spot_check(eng):
    """configs[3] shape (m=32, p=16, r=16), 592 draws (4 per SM) x T=100: SquareRootFilter logp + grad against the
    oracle on three of them (the oracle needs seconds per series at this shape)."""
    from pymc_experimental_b200 import synth

    cfg = synth.varmax(B=592, n=100)
    dcfg = _dev(cfg)
    logp, grads, status = eng.logp_grad("cholesky", dcfg["y"], *[dcfg[k] for k in NAMES])
    assert eng.last_path == "sqrt_warp_dmma+adjoint_4warp"
    assert int(status.abs().sum()) == 0
    _spot_check("cholesky", cfg, logp.cpu().numpy(), {k: v.cpu().numpy() for k, v in grads.items()},
                np.array([0, 295, 591]), rtol=1e-9)


def test_c3_conditional_posterior_draw_full_size(eng):
    """Row f-3 at the C3 size (8 192 draws x T=240, m=15): one draw per (draw, step) from the smoother's output; the
    Mahalanobis norm of every draw equals the norm of its variates, (x - mu)^T cov^-1 (x - mu) == |z|^2 -- an identity
    that holds for any square root of cov -- and the empirical moments over the batch are those of a standard normal."""
    import torch

    from pymc_experimental_b200 import synth

    cfg = synth.structural(B=8192, n=240)
    dcfg = _dev(cfg)
    ll, sa, sP, status = eng.filter_smooth("standard", dcfg["y"], *[dcfg[k] for k in NAMES])
    gen = torch.Generator(device="cuda:0").manual_seed(11)
    z = torch.randn(sa.shape, dtype=torch.float64, device="cuda:0", generator=gen)
    x, st = eng.sample_mvn(sa, sP, z)
    assert eng.last_path == "mvn_chol_subwarp"
    assert int(st.abs().sum()) == 0
    dx = (x - sa).unsqueeze(-1)
    maha = (dx * torch.linalg.solve(sP, dx)).sum((-1, -2))
    z2 = (z * z).sum(-1)
    assert float(((maha - z2).abs() / z2).max()) < 1e-6  # cond(P_s) ~ 1e5 on these draws
    assert abs(float(z.mean())) < 1e-3 and abs(float(z.var()) - 1.0) < 1e-3
