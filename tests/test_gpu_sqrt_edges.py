"""GPU edge tests of the SquareRootFilter path the verdict of round 1 asked for: run-to-run determinism of the adjoint
(its accumulators are L2 reductions), the full C4 horizon (T = 1000) against the oracle, float32, and the measured parity
margins that justify every tolerance looser than 1e-9 (profiles/r02_parity_margins.json, tools/parity_margins.py)."""

import json
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    from pymc_experimental_b200.engine import KalmanEngine

    return KalmanEngine(0)


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("shape", [(32, 16, 16), (16, 8, 8)])
def test_square_root_adjoint_is_bit_reproducible(eng, shape):
    """The adjoint's add-only accumulators (T-bar, Z-bar, sum of Psi) are fire-and-forget reductions into L2: every element
    has ONE owner lane and one addend per step, issued in step order, so the result must not depend on the run."""
    from pymc_experimental_b200 import synth

    m, p, r = shape
    cfg = synth.random_dense(148 * 2 + 5, 60, m, p, r, seed=11, missing=0.0)
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    runs = [eng.logp_grad("cholesky", *args) for _ in range(3)]
    for logp, grads, _ in runs[1:]:
        assert np.array_equal(logp, runs[0][0])
        for k in NAMES:
            assert np.array_equal(grads[k], runs[0][1][k]), k


def test_c4_full_horizon_against_the_oracle(eng):
    """VARMAX(2,0) k = 16 (m = 32, p = 16, r = 16) at the full T = 1000 of the BASELINE config: rounding accumulates over
    2000 QR sweeps.  Three series through autograd of the torch oracle."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    cfg = synth.varmax(3, 1000)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("cholesky", *args)
    assert (np.asarray(status) == 0).all()
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args)
    assert_allclose(logp, ref_logp, rtol=1e-9)
    worst = {k: _rel(grads[k], ref_g[k]) for k in NAMES}
    print("C4 T=1000 gradient errors relative to max|g|:", {k: f"{v:.1e}" for k, v in worst.items()})
    for k in NAMES:
        assert worst[k] <= 1e-9, (k, worst[k])


@pytest.mark.parametrize("shape", [(16, 8, 8), (32, 16, 16)])
@pytest.mark.parametrize("eps", [1e-3, 1e-5])
def test_square_root_nearly_dependent_columns_take_the_explicit_norm(eng, shape, eps):
    """The forward sweep's reflector scalars come from downdated column norms (panel_reflectors_la); a column that lost more
    than 2^-6 of its squared norm inside a panel must fall back to the explicit norm.  The factor P0 handed in has pairs of
    nearly parallel ROWS (= columns of the transposed update array), in the first and in a later panel, so the guard fires
    in the first steps (dlarfg's tau = 0 is exercised by the H = 0 and zero-Z-row cases of test_gpu_parity.py).  The result must match the oracle as well as the
    conditioning allows (cond(P0 factor) ~ 1/eps: bound 1e-9 / eps relative to the 1e-16 of the arithmetic, stated below)."""
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 6, 12
    cfg = synth.random_dense(B, n, m, p, r, seed=77 + m, missing=0.0)
    rng = np.random.default_rng(5)
    L = np.linalg.cholesky(cfg["P0"]) @ np.linalg.qr(rng.standard_normal((B, m, m)))[0]  # a dense factor of the same P0 (Q5)
    L[:, 1, :] = L[:, 0, :] + eps * np.abs(L[:, 0, :]).max() * rng.standard_normal((B, m))    # first panel
    L[:, 11, :] = L[:, 9, :] + eps * np.abs(L[:, 9, :]).max() * rng.standard_normal((B, m))   # second panel
    cfg["P0"] = L
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("cholesky", *args)
    assert eng.last_path == "sqrt_warp_dmma+adjoint_4warp"
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args)
    tol = max(1e-9, 1e-13 / eps**2)  # the pull-back through the factor divides by its small pivots twice
    assert_allclose(logp, ref_logp, rtol=tol)
    for k in NAMES:
        assert _rel(grads[k], ref_g[k]) <= 10 * tol, (k, _rel(grads[k], ref_g[k]))


@pytest.mark.parametrize("shape", [(5, 2, 3), (16, 8, 8), (32, 16, 16)])
def test_square_root_filter_float32(eng, shape):
    """float32 SquareRootFilter (CTA-per-series kernels) against the float64 oracle at the float32 bound of the
    north_star: 1e-3 on logp, 1e-2 on gradients (relative to max|g| of the matrix)."""
    import torch

    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    m, p, r = shape
    B, n = 6, 12
    cfg = synth.random_dense(B, n, m, p, r, seed=500 + m, missing=0.0)
    cfg["y"][2, 5] = np.nan
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args64 = [cfg["y"]] + [cfg[k] for k in NAMES]
    args32 = [torch.as_tensor(a, dtype=torch.float32, device="cuda") for a in args64]
    logp, grads, status = eng.logp_grad("cholesky", *args32)
    assert eng.last_path == "sqrt_cta" and logp.dtype == torch.float32
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args64, cov_jitter=1e-6)
    assert_allclose(logp.cpu().numpy(), ref_logp, rtol=1e-3)
    for k in NAMES:
        assert _rel(grads[k].cpu().numpy(), ref_g[k]) <= 1e-2, k
    out = eng.filter("cholesky", *args32)
    ref = oth.kalman_filter("cholesky", *[torch.as_tensor(a) for a in args64], cov_jitter=1e-6)
    assert _rel(out["filtered_states"].cpu().numpy(), ref[0].numpy()) <= 1e-3
    assert _rel(out["filtered_covariances"].cpu().numpy(), ref[3].numpy()) <= 1e-3


def test_parity_margins_are_committed_and_cover_every_loosened_tolerance():
    """Every tolerance in the GPU tests that is looser than the north_star's 1e-9 names a case in
    profiles/r02_parity_margins.json, where the measured error and the condition estimate x epsilon that explains it are
    recorded (tools/parity_margins.py regenerates the file on a GPU box)."""
    path = os.path.join(ROOT, "profiles", "r02_parity_margins.json")
    data = json.load(open(path))
    for case in ("square_root_batched", "square_root_corner_H_zero", "nile_diffuse", "missing_nondiag_H", "smoother_diffuse"):
        assert case in data, case
        rec = data[case]
        assert rec["measured_rel_error"] <= rec["tolerance_in_tests"]
        assert rec["condition_times_eps"] > 0
