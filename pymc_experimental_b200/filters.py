"""Drop-in mirrors of ``pymc_extras.statespace.filters`` backed by the CUDA engine.

Same class names, ``build_graph`` signatures, attributes and output order as the reference:

* ``BaseFilter.build_graph(data, a0, P0, c, d, T, Z, R, H, Q, mode=None, return_updates=False,
  missing_fill_value=None, cov_jitter=None)``  -- kalman_filter.py:144-238; returns
  ``[filtered_states, predicted_states, observed_states, filtered_covariances, predicted_covariances,
  observed_covariances, loglike_obs]`` (kalman_filter.py:298-308) and sets ``seq_names / non_seq_names /
  n_states / n_shocks / n_endog / mode / missing_fill_value / cov_jitter`` which ``PyMCStateSpace`` reads
  (core/statespace.py:628, 1267, 2053, 2249).
* ``KalmanSmoother.build_graph(T, R, Q, filtered_states, filtered_covariances, mode=None, cov_jitter=...)``
  -- kalman_smoother.py:66-105; returns ``(smoothed_states, smoothed_covariances)``.
* ``FILTER_FACTORY`` -- core/statespace.py:51-55.

Two calling modes:

1. **symbolic** (PyTensor installed and any input is a PyTensor variable): the graph contains one
   :class:`~pymc_experimental_b200.pytensor_ops.KalmanFilterOp` whose ``perform`` / ``L_op`` call libbkf.so, so
   ``pm.sample`` gets logp and dlogp from the GPU.  ``mode`` must be ``None`` (C/Python linker): there is
   deliberately no JAX / Numba dispatch and no CPU fallback.
2. **eager** (NumPy arrays or torch CUDA tensors): the kernels run immediately and arrays are returned.  Every
   input may carry a leading batch axis (series x draws); this is what the tests and the benchmark use, and what a
   vectorised caller (``pm.sample_posterior_predictive`` over all draws at once) would use.
"""

from __future__ import annotations

import numpy as np

from .engine import JITTER_F32, JITTER_F64, MISSING_FILL, default_engine

JITTER_DEFAULT = JITTER_F64  # utils/constants.py:20 for floatX == float64 (float32: 1e-6, see _resolve_jitter)
PARAM_NAMES = ["c", "d", "T", "Z", "R", "H", "Q"]  # kalman_filter.py:22
_STATIC_NDIM = {"c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}
OUTPUT_NAMES = ["filtered_states", "predicted_states", "observed_states", "filtered_covariances",
                "predicted_covariances", "observed_covariances", "loglike_obs"]


def _resolve_jitter(cov_jitter, symbolic, arrays):
    """``cov_jitter=None`` means the reference's dtype-dependent default (utils/constants.py:20): 1e-8 in float64,
    1e-6 in float32 -- floatX for a PyTensor graph, the dtype the engine will compute in for eager inputs (float64 for
    NumPy arrays, the tensors' own dtype for torch CUDA tensors)."""
    if cov_jitter is not None:
        return cov_jitter
    if symbolic:
        from . import pytensor_ops

        return pytensor_ops.default_jitter()
    for a in arrays:
        if type(a).__module__.startswith("torch") and getattr(a, "is_cuda", False):
            return JITTER_F32 if str(a.dtype) == "torch.float32" else JITTER_F64
    return JITTER_F64


def _is_symbolic(*xs):
    try:
        from pytensor.graph.basic import Variable
    except ImportError:
        return False
    return any(isinstance(x, Variable) for x in xs)


class BaseFilter:
    """kalman_filter.py:30-68.  ``kind`` selects the CUDA kernel family."""

    kind: str | None = None

    def __init__(self, mode=None, device: int = 0):
        self.mode = mode
        self.seq_names: list[str] = []
        self.non_seq_names: list[str] = []
        self.n_states = None
        self.n_posdef = None
        self.n_shocks = None
        self.n_endog = None
        self.missing_fill_value = None
        self.cov_jitter = None
        self.device = device

    def check_params(self, data, a0, P0, c, d, T, Z, R, H, Q):
        """kalman_filter.py:70-74."""
        return data, a0, P0, c, d, T, Z, R, H, Q

    def update(self, a, P, y, d, Z, H, all_nan_flag):
        """kalman_filter.py:413-456: the base class has no update rule."""
        raise NotImplementedError

    def build_graph(self, data, a0, P0, c, d, T, Z, R, H, Q, mode=None, return_updates=False,
                    missing_fill_value=None, cov_jitter=None, time_varying=None):
        """``time_varying`` (eager mode only, not in the reference signature): names of the inputs among
        c, d, T, Z, R, H, Q that carry a time axis.  ``None`` = decide by ndim as the reference does
        (filters/utilities.py:8-29) -- unambiguous only for unbatched inputs, so batched eager callers pass it."""
        if self.kind is None:
            raise NotImplementedError
        if missing_fill_value is None:
            missing_fill_value = MISSING_FILL  # kalman_filter.py:197-200
        data, a0, P0, c, d, T, Z, R, H, Q = self.check_params(data, a0, P0, c, d, T, Z, R, H, Q)
        symbolic = _is_symbolic(data, a0, P0, c, d, T, Z, R, H, Q)
        cov_jitter = _resolve_jitter(cov_jitter, symbolic, (data, a0, P0, c, d, T, Z, R, H, Q))
        self.mode = mode
        self.missing_fill_value = missing_fill_value
        self.cov_jitter = cov_jitter

        if symbolic:
            if mode not in (None, "FAST_RUN", "FAST_COMPILE"):
                raise NotImplementedError(
                    f"mode={mode!r}: the B200 Kalman Ops only run under PyTensor's default C/Python linker "
                    "(no JAX / Numba dispatch, no CPU fallback)")
            from . import pytensor_ops

            outs = pytensor_ops.build_filter_graph(self, data, a0, P0, c, d, T, Z, R, H, Q)
            return (outs, {}) if return_updates else outs

        # ---- eager ----
        R_shape, Z_shape = np.shape(R), np.shape(Z)
        self.n_states, self.n_shocks = int(R_shape[-2]), int(R_shape[-1])  # kalman_filter.py:206-210
        self.n_posdef = self.n_shocks
        self.n_endog = int(Z_shape[-2])
        vals = dict(zip(PARAM_NAMES, (c, d, T, Z, R, H, Q)))
        if time_varying is None:
            if np.ndim(data) != 2:
                time_varying = ()
            else:  # decide_if_x_time_varies, filters/utilities.py:8-29
                time_varying = tuple(k for k in PARAM_NAMES if np.ndim(vals[k]) == _STATIC_NDIM[k] + 1)
        self.seq_names = [k for k in PARAM_NAMES if k in time_varying]  # kalman_filter.py:212-221
        self.non_seq_names = [k for k in PARAM_NAMES if k not in time_varying]
        eng = default_engine(self.device)
        res = eng.filter(self.kind, data, a0, P0, c, d, T, Z, R, H, Q, cov_jitter=cov_jitter,
                         missing_fill_value=missing_fill_value, time_varying=tuple(self.seq_names))
        outs = [res[k] for k in OUTPUT_NAMES]
        self.status = res["status"]
        return (outs, {}) if return_updates else outs

    def logp_and_grad(self, data, a0, P0, c, d, T, Z, R, H, Q, ll_cotangent=None, missing_fill_value=None,
                      cov_jitter=None, time_varying=()):
        """Eager ``(sum_t loglike_obs, {name: d/d name})`` -- what ``model.logp_dlogp_function`` needs from the
        filter."""
        eng = default_engine(self.device)
        logp, grads, status = eng.logp_grad(self.kind, data, a0, P0, c, d, T, Z, R, H, Q, ll_cotangent=ll_cotangent,
                                            cov_jitter=cov_jitter, missing_fill_value=missing_fill_value,
                                            time_varying=tuple(time_varying))
        self.status = status
        return logp, grads


class StandardFilter(BaseFilter):
    """kalman_filter.py:544-617."""

    kind = "standard"


class UnivariateFilter(BaseFilter):
    """kalman_filter.py:753-820."""

    kind = "univariate"


class SquareRootFilter(BaseFilter):
    """kalman_filter.py:620-750.  ``P0`` is taken as a Cholesky-like factor, as in the reference."""

    kind = "cholesky"


class KalmanSmoother:
    """kalman_smoother.py:15-126."""

    def __init__(self, mode: str | None = None, device: int = 0):
        self.mode = mode
        self.cov_jitter = JITTER_DEFAULT
        self.seq_names: list[str] = []
        self.non_seq_names: list[str] = []
        self.device = device

    def build_graph(self, T, R, Q, filtered_states, filtered_covariances, mode=None, cov_jitter=None,
                    time_varying=None):
        symbolic = _is_symbolic(T, R, Q, filtered_states, filtered_covariances)
        cov_jitter = _resolve_jitter(cov_jitter, symbolic, (T, R, Q, filtered_states, filtered_covariances))
        self.mode = mode
        self.cov_jitter = cov_jitter
        if symbolic:
            if mode not in (None, "FAST_RUN", "FAST_COMPILE"):
                raise NotImplementedError(f"mode={mode!r} is not supported by the B200 Kalman Ops")
            from . import pytensor_ops

            return pytensor_ops.build_smoother_graph(self, T, R, Q, filtered_states, filtered_covariances)
        vals = {"T": T, "R": R, "Q": Q}
        if time_varying is None:
            time_varying = () if np.ndim(filtered_states) != 2 else tuple(k for k in vals if np.ndim(vals[k]) == 3)
        self.seq_names = [k for k in vals if k in time_varying]  # kalman_smoother.py:79-82
        self.non_seq_names = [k for k in vals if k not in time_varying]
        eng = default_engine(self.device)
        return eng.smooth(T, R, Q, filtered_states, filtered_covariances, cov_jitter=cov_jitter,
                          time_varying=tuple(self.seq_names))


# core/statespace.py:51-55
FILTER_FACTORY = {
    "standard": StandardFilter,
    "univariate": UnivariateFilter,
    "cholesky": SquareRootFilter,
}


def install_into_pymc_extras():
    """Make every ``PyMCStateSpace`` (SARIMAX, VARMAX, ETS, structural ...) pick up the CUDA filters:
    overrides the three ``FILTER_FACTORY`` entries and the ``KalmanSmoother`` used by
    ``pymc_extras.statespace.core.statespace`` (core/statespace.py:51-55, 253-254).  Needs pymc_extras."""
    import pymc_extras.statespace.core.statespace as ss

    ss.FILTER_FACTORY.update(FILTER_FACTORY)
    ss.KalmanSmoother = KalmanSmoother
    return ss.FILTER_FACTORY
