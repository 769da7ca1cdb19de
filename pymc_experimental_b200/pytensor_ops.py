"""PyTensor Ops that put the CUDA Kalman engine into a PyMC graph (perform + L_op over the C ABI via ctypes).

PyTensor is NOT installable in the build / GPU containers of this project, so this module is
import-guarded and deliberately thin: all numerics live behind ``KalmanEngine`` (exercised by the GPU tests through
the same ctypes calls).  The Op protocol follows the in-repo precedent ``pymc_extras/utils/spline.py``
(``BSplineBasis``: ``__props__``, ``make_node``, ``perform``, ``infer_shape``) plus ``L_op``.

Design notes
* Ops hold only picklable props (kind, jitter, fill, device); the engine handle is looked up lazily per process
  (``engine.default_engine``), so PyMC's forked chain workers each create their own CUDA context on first use.
* Only the ``loglike_obs`` cotangent is differentiable (what NUTS needs: ``SequenceMvNormal``'s logp returns it
  untouched, filters/distributions.py:446-453).  A connected cotangent on any other output raises.
* No JAX / Numba dispatch is registered on purpose (no CPU fallback): the Ops run under the default linker only.
"""

from __future__ import annotations

import numpy as np
import pytensor
import pytensor.tensor as pt
from pytensor.gradient import DisconnectedType, grad_undefined
from pytensor.graph.basic import Apply
from pytensor.graph.op import Op

from .engine import PARAM_NAMES, default_engine

OUTPUT_NAMES = ["filtered_states", "predicted_states", "observed_states", "filtered_covariances",
                "predicted_covariances", "observed_covariances", "loglike_obs"]
_ND = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


class KalmanFilterOp(Op):
    """data, a0, P0, c, d, T, Z, R, H, Q -> the seven outputs of ``BaseFilter.build_graph``."""

    __props__ = ("kind", "cov_jitter", "missing_fill_value", "device", "time_varying")

    def __init__(self, kind, cov_jitter, missing_fill_value, device=0, time_varying=()):
        self.kind, self.cov_jitter, self.missing_fill_value, self.device = kind, float(cov_jitter), float(
            missing_fill_value), int(device)
        self.time_varying = tuple(time_varying)  # names with a leading time axis (filters/utilities.py:8-47)

    def make_node(self, data, *params):
        data = pt.as_tensor_variable(data)
        params = [pt.as_tensor_variable(x) for x in params]
        if data.ndim != 2:
            raise ValueError("data must be (time, observed_state)")
        for name, x in zip(PARAM_NAMES, params):
            if x.ndim != _ND[name] + int(name in self.time_varying):
                raise ValueError(f"{name} has {x.ndim} dims, expected {_ND[name] + int(name in self.time_varying)}")
        dtype = pytensor.config.floatX
        n = data.type.shape[0]
        m, r = params[6].type.shape[-2:]  # R
        p = params[5].type.shape[-2]  # Z
        po = 1 if self.kind == "univariate" else p
        shapes = [(n, m), (n, m), (n, po), (n, m, m), (n, m, m), (n, po, po), (n,)]
        outs = [pt.tensor(dtype=dtype, shape=s) for s in shapes]
        return Apply(self, [data, *params], outs)

    def perform(self, node, inputs, output_storage, params=None):
        data, *prm = inputs
        res = default_engine(self.device).filter(
            self.kind, data, *prm, cov_jitter=self.cov_jitter, missing_fill_value=self.missing_fill_value,
            dtype=node.outputs[0].dtype, time_varying=self.time_varying)
        for store, name in zip(output_storage, OUTPUT_NAMES):
            store[0] = res[name]

    def infer_shape(self, fgraph, node, ins_shapes):
        (n, p) = ins_shapes[0]
        m = ins_shapes[1][0]  # a0
        po = 1 if self.kind == "univariate" else p
        return [(n, m), (n, m), (n, po), (n, m, m), (n, m, m), (n, po, po), (n,)]

    def connection_pattern(self, node):
        # data has no gradient; every parameter is connected to every output
        return [[False] * 7] + [[True] * 7 for _ in range(9)]

    def L_op(self, inputs, outputs, output_grads):
        for name, g in zip(OUTPUT_NAMES[:-1], output_grads[:-1]):
            if not isinstance(g.type, DisconnectedType):
                raise NotImplementedError(
                    f"gradient through '{name}' is not implemented by the CUDA engine (only loglike_obs is)")
        g_ll = output_grads[-1]
        if isinstance(g_ll.type, DisconnectedType):
            return [grad_undefined(self, 0, inputs[0])] + [pt.zeros_like(x) for x in inputs[1:]]
        grads = KalmanGradOp(self.kind, self.cov_jitter, self.missing_fill_value, self.device,
                             self.time_varying)(*inputs, g_ll)
        return [grad_undefined(self, 0, inputs[0], "no gradient w.r.t. the observed data"), *grads]


class KalmanGradOp(Op):
    """data, a0..Q, ll_cotangent -> d(sum_t cot_t ll_t)/d{a0,P0,c,d,T,Z,R,H,Q}  (bkf_filter_logp_grad)."""

    __props__ = ("kind", "cov_jitter", "missing_fill_value", "device", "time_varying")

    def __init__(self, kind, cov_jitter, missing_fill_value, device=0, time_varying=()):
        self.kind, self.cov_jitter, self.missing_fill_value, self.device = kind, float(cov_jitter), float(
            missing_fill_value), int(device)
        self.time_varying = tuple(time_varying)

    def make_node(self, data, *rest):
        ins = [pt.as_tensor_variable(x) for x in (data, *rest)]
        outs = [x.type() for x in ins[1:10]]
        return Apply(self, ins, outs)

    def perform(self, node, inputs, output_storage, params=None):
        data, *prm, cot = inputs
        _, grads, _ = default_engine(self.device).logp_grad(
            self.kind, data, *prm, ll_cotangent=np.ascontiguousarray(cot), cov_jitter=self.cov_jitter,
            missing_fill_value=self.missing_fill_value, dtype=node.outputs[0].dtype,
            time_varying=self.time_varying)
        for store, name in zip(output_storage, PARAM_NAMES):
            store[0] = grads[name]

    def infer_shape(self, fgraph, node, ins_shapes):
        return list(ins_shapes[1:10])


class KalmanSmootherOp(Op):
    """T, R, Q, filtered_states, filtered_covariances -> smoothed_states, smoothed_covariances."""

    __props__ = ("cov_jitter", "device", "time_varying")

    def __init__(self, cov_jitter, device=0, time_varying=()):
        self.cov_jitter, self.device, self.time_varying = float(cov_jitter), int(device), tuple(time_varying)

    def make_node(self, T, R, Q, fa, fP):
        ins = [pt.as_tensor_variable(x) for x in (T, R, Q, fa, fP)]
        for name, x in zip(("T", "R", "Q"), ins[:3]):
            if x.ndim != 2 + int(name in self.time_varying):
                raise ValueError(f"{name} has {x.ndim} dims, expected {2 + int(name in self.time_varying)}")
        return Apply(self, ins, [ins[3].type(), ins[4].type()])

    def perform(self, node, inputs, output_storage, params=None):
        sa, sP = default_engine(self.device).smooth(*inputs, cov_jitter=self.cov_jitter, dtype=node.outputs[0].dtype,
                                                    time_varying=self.time_varying)
        output_storage[0][0], output_storage[1][0] = sa, sP

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[3], ins_shapes[4]]


class StationaryInitOp(Op):
    """c, T, R, Q -> x0 = (I - T)^-1 c, P0 = solve_discrete_lyapunov(T, R Q R^T)  (bkf_stationary_init; models/SARIMAX.py:
    369-381).  ``L_op`` applies :class:`StationaryInitGradOp` (bkf_stationary_init_grad), i.e. the pull-backs of
    PyTensor's ``Solve`` and ``SolveDiscreteLyapunov`` Ops."""

    __props__ = ("device",)

    def __init__(self, device=0):
        self.device = int(device)

    def make_node(self, c, T, R, Q):
        ins = [pt.as_tensor_variable(x) for x in (c, T, R, Q)]
        if [x.ndim for x in ins] != [1, 2, 2, 2]:
            raise ValueError("c, T, R, Q must be a vector and three matrices")
        return Apply(self, ins, [ins[0].type(), ins[1].type()])

    def perform(self, node, inputs, output_storage, params=None):
        a0, P0, status = default_engine(self.device).stationary_init(*inputs, dtype=node.outputs[0].dtype)
        if int(status) & 8:
            raise np.linalg.LinAlgError("stationary initialisation: the transition matrix is not stationary")
        output_storage[0][0], output_storage[1][0] = a0, P0

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[0], ins_shapes[1]]

    def L_op(self, inputs, outputs, output_grads):
        c, T, R, Q = inputs
        gb = [pt.zeros_like(o) if isinstance(g.type, DisconnectedType) else g for o, g in zip(outputs, output_grads)]
        return StationaryInitGradOp(self.device)(T, R, Q, outputs[0], outputs[1], *gb)


class StationaryInitGradOp(Op):
    """T, R, Q, x0, P0, x0_bar, P0_bar -> c_bar, T_bar, R_bar, Q_bar."""

    __props__ = ("device",)

    def __init__(self, device=0):
        self.device = int(device)

    def make_node(self, T, R, Q, a0, P0, a0_bar, P0_bar):
        ins = [pt.as_tensor_variable(x) for x in (T, R, Q, a0, P0, a0_bar, P0_bar)]
        return Apply(self, ins, [ins[3].type(), ins[0].type(), ins[1].type(), ins[2].type()])

    def perform(self, node, inputs, output_storage, params=None):
        g = default_engine(self.device).stationary_init_grad(*inputs, dtype=node.outputs[0].dtype)
        for store, name in zip(output_storage, ("c", "T", "R", "Q")):
            store[0] = g[name]

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[3], ins_shapes[0], ins_shapes[1], ins_shapes[2]]


def build_filter_graph(flt, data, a0, P0, c, d, T, Z, R, H, Q):
    """Symbolic branch of ``BaseFilter.build_graph``: one Op node, outputs named / shaped as the reference does
    (kalman_filter.py:271-296)."""
    R_shape, Z_shape = R.type.shape, Z.type.shape
    flt.n_states, flt.n_shocks = R_shape[-2:]
    flt.n_posdef = flt.n_shocks
    flt.n_endog = Z_shape[-2]
    names = ["c", "d", "T", "Z", "R", "H", "Q"]
    nd = dict(zip(names, (c.ndim, d.ndim, T.ndim, Z.ndim, R.ndim, H.ndim, Q.ndim)))
    flt.seq_names = [k for k in names if nd[k] == _ND[k] + 1]  # decide_if_x_time_varies, filters/utilities.py:8-29
    flt.non_seq_names = [k for k in names if k not in flt.seq_names]
    op = KalmanFilterOp(flt.kind, flt.cov_jitter, flt.missing_fill_value, flt.device, tuple(flt.seq_names))
    outs = op(data, a0, P0, c, d, T, Z, R, H, Q)
    for o, name in zip(outs, OUTPUT_NAMES):
        o.name = name
    return list(outs)


def build_smoother_graph(sm, T, R, Q, filtered_states, filtered_covariances):
    nd = {"T": T.ndim, "R": R.ndim, "Q": Q.ndim}
    sm.seq_names = [k for k in nd if nd[k] == 3]
    sm.non_seq_names = [k for k in nd if nd[k] != 3]
    sa, sP = KalmanSmootherOp(sm.cov_jitter, sm.device, tuple(sm.seq_names))(
        T, R, Q, filtered_states, filtered_covariances)
    sa.name, sP.name = "smoothed_states", "smoothed_covariances"
    return sa, sP
