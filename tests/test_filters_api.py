"""The host-side mirror of pymc_extras.statespace.filters (pymc_experimental_b200/filters.py): these tests read like
the reference's own tests/statespace/test_kalman_filter.py, with the oracle standing in for the compiled PyTensor
function."""

import numpy as np
import pytest
from numpy.testing import assert_allclose

from pymc_experimental_b200.filters import (
    FILTER_FACTORY,
    BaseFilter,
    KalmanSmoother,
    SquareRootFilter,
    StandardFilter,
    UnivariateFilter,
)

FILTERS = [StandardFilter, SquareRootFilter, UnivariateFilter]
IDS = ["StandardFilter", "CholeskyFilter", "UnivariateFilter"]
KIND = {StandardFilter: "standard", SquareRootFilter: "cholesky", UnivariateFilter: "univariate"}


def make_test_inputs(p, m, r, n, rng, missing_data=None, H_is_zero=False):
    """tests/statespace/utilities/test_helpers.py:86-111."""
    data = np.arange(n * p, dtype=float).reshape(-1, p)
    if missing_data is not None:
        data[rng.choice(n, missing_data, replace=False)] = np.nan
    T = np.eye(m, k=-1)
    T[0, :] = 1 / m
    return (data, np.zeros(m), np.eye(m), np.zeros(m), np.zeros(p), T, np.eye(m)[:p, :], np.eye(m)[:, :r],
            np.zeros((p, p)) if H_is_zero else np.eye(p), np.eye(r))


def run(cls, inputs):
    kf, ks = cls(), KalmanSmoother()
    outs = kf.build_graph(*inputs)
    T, R, Q = inputs[5], inputs[7], inputs[9]
    sa, sP = ks.build_graph(T, R, Q, outs[0], outs[3])
    return kf, outs, sa, sP


def test_base_class_update_raises():
    """test_kalman_filter.py:61-65."""
    with pytest.raises(NotImplementedError):
        BaseFilter().update(*[None] * 7)
    with pytest.raises(NotImplementedError):
        BaseFilter().build_graph(*[None] * 10)


def test_factory_keys_match_reference():
    """core/statespace.py:51-55."""
    assert set(FILTER_FACTORY) == {"standard", "univariate", "cholesky"}
    assert FILTER_FACTORY["cholesky"] is SquareRootFilter


def test_symbolic_layer_is_import_guarded():
    try:
        import pytensor  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            import pymc_experimental_b200.pytensor_ops  # noqa: F401
    else:
        import pymc_experimental_b200.pytensor_ops  # noqa: F401


@pytest.mark.gpu
@pytest.mark.parametrize("cls", FILTERS, ids=IDS)
@pytest.mark.parametrize("shape", [(1, 1, 1, 10), (1, 2, 2, 10), (1, 5, 2, 10), (1, 5, 1, 10), (5, 5, 1, 10)])
def test_output_shapes_and_attributes(cls, shape, rng):
    """test_kalman_filter.py:69-195 (shape tests) + the attributes PyMCStateSpace reads."""
    p, m, r, n = shape
    kf, outs, sa, sP = run(cls, make_test_inputs(p, m, r, n, rng))
    po = 1 if cls is UnivariateFilter else p
    expected = [(n, m), (n, m), (n, po), (n, m, m), (n, m, m), (n, po, po), (n,)]
    assert [o.shape for o in outs] == expected
    assert sa.shape == (n, m) and sP.shape == (n, m, m)
    assert (kf.n_states, kf.n_shocks, kf.n_endog) == (m, r, p)
    assert kf.seq_names == [] and kf.non_seq_names == ["c", "d", "T", "Z", "R", "H", "Q"]
    assert kf.missing_fill_value == -9999.0 and kf.cov_jitter == 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("cls", FILTERS, ids=IDS)
@pytest.mark.parametrize("p", [1, 5])
def test_missing_data(cls, p, rng):
    """test_kalman_filter.py:198-211."""
    m, r, n = 5, 1, 10
    inputs = make_test_inputs(p, m, r, n, rng, missing_data=1)
    _, outs, sa, sP = run(cls, inputs)
    assert outs[0].shape == (n, m) and sa.shape == (n, m)
    assert np.isfinite(outs[6]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("cls", FILTERS, ids=IDS)
def test_last_smoother_is_last_filtered(cls, rng):
    """test_kalman_filter.py:214-224."""
    _, outs, sa, sP = run(cls, make_test_inputs(1, 5, 1, 10, rng))
    assert_allclose(outs[0][-1], sa[-1])
    assert_allclose(outs[3][-1], sP[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("cls", FILTERS, ids=IDS)
@pytest.mark.parametrize("shape", [(1, 5, 2, 10), (5, 5, 1, 10), (3, 6, 2, 12)])
def test_outputs_match_oracle(cls, shape, rng):
    from oracle import kalman_np as onp

    p, m, r, n = shape
    inputs = make_test_inputs(p, m, r, n, rng)
    _, outs, sa, sP = run(cls, inputs)
    ref = onp.kalman_filter(KIND[cls], *inputs)
    for o, e in zip(outs, ref):
        assert_allclose(o, e.reshape(o.shape), rtol=1e-9, atol=1e-9 * max(1.0, np.abs(e).max()))
    ra, rP = onp.kalman_smoother(inputs[5], inputs[7], inputs[9], ref[0], ref[3])
    assert_allclose(sa, ra, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(ra).max()))
    assert_allclose(sP, rP, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(rP).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("cls", [StandardFilter, SquareRootFilter], ids=IDS[:2])
@pytest.mark.parametrize("n_missing", [0, 5])
@pytest.mark.parametrize("obs_noise", [True, False])
def test_all_covariance_matrices_are_PSD(cls, n_missing, obs_noise):
    """test_kalman_filter.py:271-301 on the Nile inputs (golden fixture = reference's test inputs)."""
    from conftest import load_golden

    _, ins, _ = load_golden(f"nile_standard_miss{n_missing}")
    P0 = np.linalg.cholesky(ins["P0"]) if cls is SquareRootFilter else ins["P0"]
    inputs = (ins["y"], ins["a0"], P0, ins["c"], ins["d"], ins["T"], ins["Z"], ins["R"], ins["H"] * int(obs_noise),
              ins["Q"])
    _, outs, sa, sP = run(cls, inputs)
    for cov in (outs[3], outs[4], sP):
        w = np.linalg.eigvals(cov)
        assert (w > 0).all()
        assert_allclose(cov, np.swapaxes(cov, -2, -1), rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_eager_batched_build_graph_and_logp_grad(rng):
    from oracle import kalman_torch as oth
    from pymc_experimental_b200 import synth

    cfg = synth.ets(B=33, n=40)
    names = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
    kf = StandardFilter()
    outs = kf.build_graph(cfg["y"], *[cfg[k] for k in names])
    assert outs[6].shape == (33, 40)
    logp, grads = kf.logp_and_grad(cfg["y"], *[cfg[k] for k in names])
    ref_logp, ref_g = oth.logp_and_grad("standard", cfg["y"], *[cfg[k] for k in names])
    assert_allclose(outs[6].sum(-1), ref_logp, rtol=1e-9)
    assert_allclose(logp, ref_logp, rtol=1e-9)
    for k in names:
        assert_allclose(grads[k], ref_g[k], rtol=0, atol=1e-9 * np.abs(ref_g[k]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("cls", [StandardFilter, UnivariateFilter], ids=["StandardFilter", "UnivariateFilter"])
def test_time_varying_inputs_are_detected_by_ndim_like_the_reference(cls, rng):
    """test_kalman_filter.py:151-166 (time-varying matrices) + filters/utilities.py:8-47: inputs with one extra
    leading (time) dimension become scan sequences; seq_names / non_seq_names are what PyMCStateSpace reads
    (core/statespace.py:628, 1267, 2053, 2249)."""
    from oracle import kalman_np as onp

    p, m, r, n = 1, 5, 2, 10
    data, a0, P0, c, d, T, Z, R, H, Q = make_test_inputs(p, m, r, n, rng)
    Zt = np.repeat(Z[None], n, axis=0) * (1.0 + 0.1 * np.arange(n))[:, None, None]  # regression-style Z[t]
    Qt = np.repeat(Q[None], n, axis=0) * (1.0 + 0.05 * np.arange(n))[:, None, None]
    kf, ks = cls(), KalmanSmoother()
    outs = kf.build_graph(data, a0, P0, c, d, T, Zt, R, H, Qt)
    assert kf.seq_names == ["Z", "Q"] and kf.non_seq_names == ["c", "d", "T", "R", "H"]
    ref = onp.kalman_filter(KIND[cls], data, a0, P0, c, d, T, Zt, R, H, Qt)
    for o, e in zip(outs, ref):
        assert_allclose(o, e.reshape(o.shape), rtol=1e-9, atol=1e-9 * max(1.0, np.abs(e).max()))
    sa, sP = ks.build_graph(T, R, Qt, outs[0], outs[3])
    assert ks.seq_names == ["Q"] and ks.non_seq_names == ["T", "R"]
    ra, rP = onp.kalman_smoother(T, R, Qt, ref[0], ref[3])
    assert_allclose(sa, ra, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(ra).max()))
    assert_allclose(sP, rP, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(rP).max()))


@pytest.mark.gpu
def test_solve_discrete_lyapunov_mirror(rng):
    """pymc_experimental_b200.stationary.solve_discrete_lyapunov against scipy (what the reference's "bilinear" method
    calls), unbatched and batched, and the non-stationary error."""
    import scipy.linalg

    from pymc_experimental_b200.stationary import solve_discrete_lyapunov, stationary_initialization

    m = 5
    A = rng.standard_normal((m, m))
    A *= 0.8 / np.abs(np.linalg.eigvals(A)).max()
    L = rng.standard_normal((m, m))
    Q = L @ L.T
    X = solve_discrete_lyapunov(A, Q)
    assert_allclose(X, scipy.linalg.solve_discrete_lyapunov(A, Q), rtol=1e-9, atol=1e-11)
    Xb = solve_discrete_lyapunov(np.stack([A, 0.5 * A]), np.stack([Q, Q]))
    assert_allclose(Xb[1], scipy.linalg.solve_discrete_lyapunov(0.5 * A, Q), rtol=1e-9, atol=1e-11)
    c = rng.standard_normal(m)
    x0, P0 = stationary_initialization(c, A, np.eye(m), Q)
    assert_allclose(x0, np.linalg.solve(np.eye(m) - A, c), rtol=1e-9)
    with pytest.raises(np.linalg.LinAlgError):
        solve_discrete_lyapunov(np.eye(m), Q)
