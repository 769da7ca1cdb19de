"""theta -> entries programs (pymc_experimental_b200/theta_maps.py): the NumPy restatement that checks the device path is
itself checked here -- against the formulas of the reference's models written out by hand and against finite differences."""
import numpy as np
import pytest

from pymc_experimental_b200 import theta_maps
from pymc_experimental_b200.engine import FN_COS, FN_EXP, FN_ID, FN_ONE_MINUS, FN_SIN, FN_SQUARE


def test_sarima_program_is_the_multiplied_out_lag_polynomial():
    # models/SARIMAX.py:404-458: (1 - phi L)(1 - Phi L^12) and (1 + theta L)(1 + Theta L^12)
    positions, terms, ntheta = theta_maps.sarima_111_101_12()
    rng = np.random.default_rng(0)
    th = rng.uniform(-0.9, 0.9, (11, ntheta))
    u = theta_maps.evaluate(terms, th, len(positions))
    phi, Phi, tma, Tma, sig = th.T
    np.testing.assert_allclose(u, np.column_stack([phi, Phi, -phi * Phi, tma, Tma, tma * Tma, sig**2]), rtol=1e-15)
    assert [p[0] for p in positions] == ["T", "T", "T", "R", "R", "R", "Q"]


def test_cycle_program_is_a_damped_rotation():
    # models/structural.py:1220-1237
    positions, terms, ntheta = theta_maps.structural_cycle(first_state=13, shock=3)
    th = np.array([[0.9, 0.3, -1.0], [0.5, 1.2, 0.2]])
    u = theta_maps.evaluate(terms, th, len(positions))
    rho, lam, ls = th.T
    np.testing.assert_allclose(u[:, :4], np.column_stack([rho * np.cos(lam), rho * np.sin(lam), -rho * np.sin(lam),
                                                          rho * np.cos(lam)]), rtol=1e-15)
    np.testing.assert_allclose(u[:, 4], np.exp(2 * ls), rtol=1e-15)
    np.testing.assert_allclose(u[:, 5], np.exp(2 * ls), rtol=1e-15)


@pytest.mark.parametrize("builder", [theta_maps.sarima_111_101_12, lambda: theta_maps.structural_cycle(2, 1),
                                     lambda: theta_maps.structural_cycle(2, 1, log_sigma=False), theta_maps.ets_aan,
                                     lambda: theta_maps.ets_aan(damped=False)])
def test_jacobian_matches_central_differences(builder):
    positions, terms, ntheta = builder()
    terms = terms + [(0, 0.3, [(0, FN_COS), (ntheta - 1, FN_ONE_MINUS), (0, FN_SIN)]), (1, 2.0, [(0, FN_EXP), (0, FN_SQUARE)]),
                     (1, -1.5, [(ntheta - 1, FN_ID)])]
    rng = np.random.default_rng(3)
    th = rng.uniform(0.1, 0.9, (5, ntheta))
    J = theta_maps.jacobian(terms, th, len(positions))
    h = 1e-6
    for j in range(ntheta):
        e = np.zeros(ntheta)
        e[j] = h
        fd = (theta_maps.evaluate(terms, th + e, len(positions)) - theta_maps.evaluate(terms, th - e, len(positions))) / (2 * h)
        np.testing.assert_allclose(J[:, :, j], fd, rtol=1e-7, atol=1e-9)
