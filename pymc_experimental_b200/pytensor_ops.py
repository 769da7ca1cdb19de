"""PyTensor Ops that put the CUDA Kalman engine into a PyMC graph (``perform`` + ``L_op`` over the C ABI via ctypes).

The Op protocol follows the in-repo precedent ``pymc_extras/utils/spline.py`` (``BSplineBasis``: ``__props__``,
``make_node``, ``perform``, ``infer_shape``) plus ``L_op`` / ``connection_pattern`` and the ``_vectorize_node`` dispatch
of ``pytensor.graph.replace``.  PyTensor is not installable in this project's containers, so the
layer is exercised against the protocol stub in ``tests/stubs/pytensor`` (tests/test_pytensor_ops.py; on the GPU with
the real engine, against the golden vectors).

What a graph built by ``build_filter_graph`` contains
* :class:`KalmanLoglikeOp` -- ``data, a0, P0, c, d, T, Z, R, H, Q -> loglike_obs, ticket``.  This is the node a PyMC
  model's logp hangs on (``SequenceMvNormal``'s logp returns ``loglike_obs`` untouched, filters/distributions.py:446-453).
  ``perform`` runs the forward sweep ONCE (``bkf_filter_forward_saved``), copies back only ``loglike_obs`` and leaves the
  sweep's record on the device; ``ticket`` (an int64 scalar) names that record.
* :class:`KalmanGradOp` -- applied by ``KalmanLoglikeOp.L_op`` to ``(inputs..., ticket, cotangent)``.  ``perform`` runs only
  the adjoint sweep on the saved record (``bkf_filter_backward_saved``): one evaluation of logp and dlogp is two kernel
  launches.  If the record is gone (something else ran on the engine in between) it recomputes with the fused call.
* :class:`KalmanFilterOp` -- the six state / covariance outputs of ``BaseFilter.build_graph``.  PyTensor prunes it when
  only the log-likelihood is used (``pm.sample``), and prunes :class:`KalmanLoglikeOp` when only the states are
  (``sample_conditional_posterior``), so neither path copies back arrays nobody reads.
* Every Op takes a leading batch axis on any subset of its inputs (``batch_ndim = 1``); ``vectorize_graph`` maps draws
  onto that axis through the registered ``_vectorize_node`` functions (no ``Blockwise`` Python loop).

Design notes
* Ops hold only picklable props; the engine handle is looked up lazily per process (``engine.default_engine``).  Forked
  workers cannot use CUDA: sample with ``mp_ctx="spawn"`` (the engine raises a clear error otherwise).
* Only the ``loglike_obs`` cotangent is differentiable.  The observed data has no gradient and is declared disconnected.
* No JAX / Numba dispatch is registered on purpose (no CPU fallback): the Ops run under the default linker only.
"""

from __future__ import annotations

import numpy as np
import pytensor
import pytensor.tensor as pt
from pytensor.gradient import DisconnectedType, disconnected_type
from pytensor.graph.basic import Apply
from pytensor.graph.op import Op
from pytensor.graph.replace import _vectorize_node

from .engine import JITTER_F32, JITTER_F64, PARAM_NAMES, default_engine

STATE_OUTPUT_NAMES = ["filtered_states", "predicted_states", "observed_states", "filtered_covariances",
                      "predicted_covariances", "observed_covariances"]
OUTPUT_NAMES = STATE_OUTPUT_NAMES + ["loglike_obs"]
_ND = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


def default_jitter():
    """utils/constants.py:20: 1e-8 for floatX == float64, 1e-6 for float32."""
    return JITTER_F64 if pytensor.config.floatX == "float64" else JITTER_F32


class _FilterOpBase(Op):
    """Props and input checking shared by the three filter Ops.  ``batched``: one flag per input (data, a0 .. Q) saying
    whether it carries the leading batch axis; all-False is the plain (reference) signature."""

    __props__ = ("kind", "cov_jitter", "missing_fill_value", "device", "time_varying", "batched")

    def __init__(self, kind, cov_jitter, missing_fill_value, device=0, time_varying=(), batched=None):
        self.kind, self.cov_jitter, self.missing_fill_value, self.device = kind, float(cov_jitter), float(
            missing_fill_value), int(device)
        self.time_varying = tuple(time_varying)  # names with a leading time axis (filters/utilities.py:8-47)
        self.batched = tuple(bool(x) for x in batched) if batched is not None else (False,) * 10

    @property
    def batch_ndim(self):
        return int(any(self.batched))

    def _check_inputs(self, data, params):
        data = pt.as_tensor_variable(data)
        params = [pt.as_tensor_variable(x) for x in params]
        if len(params) != 9:
            raise ValueError("expected data and the nine state-space inputs a0, P0, c, d, T, Z, R, H, Q")
        if data.ndim != 2 + int(self.batched[0]):
            raise ValueError("data must be (time, observed_state)" + (" with a leading batch axis" if self.batched[0] else ""))
        for name, x, bt in zip(PARAM_NAMES, params, self.batched[1:]):
            want = _ND[name] + int(name in self.time_varying) + int(bt)
            if x.ndim != want:
                raise ValueError(f"{name} has {x.ndim} dims, expected {want}")
        return data, params

    def _static_dims(self, data, params):
        n = data.type.shape[-2]
        m, r = params[6].type.shape[-2:]  # R
        p = params[5].type.shape[-2]  # Z
        B = None
        for x, bt in zip([data, *params], self.batched):
            if bt and x.type.shape[0] is not None:
                B = x.type.shape[0]
        return B, n, m, p, r

    def _lead(self, B):
        return (B,) if self.batch_ndim else ()

    def _engine_kwargs(self, node):
        return dict(cov_jitter=self.cov_jitter, missing_fill_value=self.missing_fill_value,
                    dtype=node.outputs[0].dtype, time_varying=self.time_varying)


class KalmanLoglikeOp(_FilterOpBase):
    """data, a0, P0, c, d, T, Z, R, H, Q -> loglike_obs [(B,) n], ticket (int64 scalar naming the saved forward sweep)."""

    def make_node(self, data, *params):
        data, params = self._check_inputs(data, params)
        B, n, m, p, r = self._static_dims(data, params)
        ll = pt.tensor(dtype=pytensor.config.floatX, shape=self._lead(B) + (n,))
        ticket = pt.tensor(dtype="int64", shape=())
        return Apply(self, [data, *params], [ll, ticket])

    def perform(self, node, inputs, output_storage, params=None):
        data, *prm = inputs
        ll, _, _, ticket = default_engine(self.device).forward_saved(self.kind, data, *prm, **self._engine_kwargs(node))
        output_storage[0][0] = np.asarray(ll)
        output_storage[1][0] = np.asarray(ticket, dtype="int64")

    def infer_shape(self, fgraph, node, ins_shapes):
        n = ins_shapes[0][-2]
        lead = ()
        for shp, bt in zip(ins_shapes, self.batched):
            if bt:
                lead = (shp[0],)
        return [lead + (n,), ()]

    def connection_pattern(self, node):
        # the data has no gradient; every parameter is connected to loglike_obs; nothing is connected to the ticket
        return [[False, False]] + [[True, False] for _ in range(9)]

    def L_op(self, inputs, outputs, output_grads):
        g_ll = output_grads[0]
        if isinstance(g_ll.type, DisconnectedType):
            return [disconnected_type()] + [pt.zeros_like(x) for x in inputs[1:]]
        grads = KalmanGradOp(self.kind, self.cov_jitter, self.missing_fill_value, self.device, self.time_varying,
                             self.batched)(*inputs, outputs[1], g_ll)
        return [disconnected_type(), *grads]


class KalmanGradOp(_FilterOpBase):
    """data, a0..Q, ticket, ll_cotangent -> d(sum_t cot_t ll_t)/d{a0,P0,c,d,T,Z,R,H,Q}.  Adjoint sweep on the record the
    ticket names (``bkf_filter_backward_saved``); the fused ``bkf_filter_logp_grad`` when that record is gone."""

    def make_node(self, data, *rest):
        if len(rest) != 11:
            raise ValueError("expected data, the nine state-space inputs, the ticket and the log-likelihood cotangent")
        data, params = self._check_inputs(data, rest[:9])
        ticket, cot = pt.as_tensor_variable(rest[9]), pt.as_tensor_variable(rest[10])
        # a gradient has the shape of its input; with a batch axis on the Op, shared inputs get per-draw gradients
        # summed over the batch (the cotangent of a broadcast is the sum)
        outs = [x.type() for x in params]
        return Apply(self, [data, *params, ticket, cot], outs)

    def perform(self, node, inputs, output_storage, params=None):
        data, *prm, ticket, cot = inputs
        eng = default_engine(self.device)
        cot = np.ascontiguousarray(cot)
        grads = eng.backward_saved(int(ticket), ll_cotangent=cot)
        if grads is None:
            _, grads, _ = eng.logp_grad(self.kind, data, *prm, ll_cotangent=cot, **self._engine_kwargs(node))
        for store, name, bt in zip(output_storage, PARAM_NAMES, self.batched[1:]):
            g = np.asarray(grads[name])
            if self.batch_ndim and not bt:
                g = g.sum(axis=0)
            store[0] = g

    def infer_shape(self, fgraph, node, ins_shapes):
        return list(ins_shapes[1:10])


class KalmanFilterOp(_FilterOpBase):
    """data, a0, P0, c, d, T, Z, R, H, Q -> the six state / covariance outputs of ``BaseFilter.build_graph``
    (kalman_filter.py:298-308 without ``loglike_obs``, which :class:`KalmanLoglikeOp` provides)."""

    def _shapes(self, B, n, m, p):
        po = 1 if self.kind == "univariate" else p
        lead = self._lead(B)
        return [lead + s for s in ((n, m), (n, m), (n, po), (n, m, m), (n, m, m), (n, po, po))]

    def make_node(self, data, *params):
        data, params = self._check_inputs(data, params)
        B, n, m, p, r = self._static_dims(data, params)
        outs = [pt.tensor(dtype=pytensor.config.floatX, shape=s) for s in self._shapes(B, n, m, p)]
        return Apply(self, [data, *params], outs)

    def perform(self, node, inputs, output_storage, params=None):
        data, *prm = inputs
        res = default_engine(self.device).filter(self.kind, data, *prm, outputs=tuple(STATE_OUTPUT_NAMES),
                                                 **self._engine_kwargs(node))
        for store, name in zip(output_storage, STATE_OUTPUT_NAMES):
            store[0] = np.asarray(res[name])

    def infer_shape(self, fgraph, node, ins_shapes):
        n, p = ins_shapes[0][-2:]
        m = ins_shapes[1][-1]  # a0
        B = None
        for shp, bt in zip(ins_shapes, self.batched):
            if bt:
                B = shp[0]
        return self._shapes(B, n, m, p)

    def L_op(self, inputs, outputs, output_grads):
        for name, g in zip(STATE_OUTPUT_NAMES, output_grads):
            if not isinstance(g.type, DisconnectedType):
                raise NotImplementedError(
                    f"gradient through '{name}' is not implemented by the CUDA engine (only loglike_obs is)")
        return [disconnected_type() for _ in inputs]

    def connection_pattern(self, node):
        return [[False] * 6 for _ in range(10)]


class KalmanSmootherOp(Op):
    """T, R, Q, filtered_states, filtered_covariances -> smoothed_states, smoothed_covariances.  ``batched``: one flag
    per input."""

    __props__ = ("cov_jitter", "device", "time_varying", "batched")

    def __init__(self, cov_jitter, device=0, time_varying=(), batched=None):
        self.cov_jitter, self.device, self.time_varying = float(cov_jitter), int(device), tuple(time_varying)
        self.batched = tuple(bool(x) for x in batched) if batched is not None else (False,) * 5

    def make_node(self, T, R, Q, fa, fP):
        ins = [pt.as_tensor_variable(x) for x in (T, R, Q, fa, fP)]
        for name, x, bt in zip(("T", "R", "Q"), ins[:3], self.batched[:3]):
            want = 2 + int(name in self.time_varying) + int(bt)
            if x.ndim != want:
                raise ValueError(f"{name} has {x.ndim} dims, expected {want}")
        if any(self.batched) and not (self.batched[3] and self.batched[4]):
            raise ValueError("a batched smoother needs batched filtered states and covariances")
        return Apply(self, ins, [ins[3].type(), ins[4].type()])

    def perform(self, node, inputs, output_storage, params=None):
        sa, sP = default_engine(self.device).smooth(*inputs, cov_jitter=self.cov_jitter, dtype=node.outputs[0].dtype,
                                                    time_varying=self.time_varying)
        output_storage[0][0], output_storage[1][0] = np.asarray(sa), np.asarray(sP)

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[3], ins_shapes[4]]


class StationaryInitOp(Op):
    """c, T, R, Q -> x0 = (I - T)^-1 c, P0 = solve_discrete_lyapunov(T, R Q R^T)  (bkf_stationary_init; models/SARIMAX.py:
    369-381).  ``L_op`` applies :class:`StationaryInitGradOp` (bkf_stationary_init_grad), i.e. the pull-backs of
    PyTensor's ``Solve`` and ``SolveDiscreteLyapunov`` Ops.  ``batched``: all four inputs carry a leading batch axis."""

    __props__ = ("device", "batched")

    def __init__(self, device=0, batched=False):
        self.device, self.batched = int(device), bool(batched)

    def make_node(self, c, T, R, Q):
        ins = [pt.as_tensor_variable(x) for x in (c, T, R, Q)]
        if [x.ndim for x in ins] != [k + int(self.batched) for k in (1, 2, 2, 2)]:
            raise ValueError("c, T, R, Q must be a vector and three matrices" + (" with a leading batch axis" if self.batched else ""))
        return Apply(self, ins, [ins[0].type(), ins[1].type()])

    def perform(self, node, inputs, output_storage, params=None):
        a0, P0, status = default_engine(self.device).stationary_init(*inputs, dtype=node.outputs[0].dtype)
        if np.any(np.asarray(status) & 8):
            raise np.linalg.LinAlgError("stationary initialisation: the transition matrix is not stationary")
        output_storage[0][0], output_storage[1][0] = np.asarray(a0), np.asarray(P0)

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[0], ins_shapes[1]]

    def L_op(self, inputs, outputs, output_grads):
        c, T, R, Q = inputs
        gb = [pt.zeros_like(o) if isinstance(g.type, DisconnectedType) else g for o, g in zip(outputs, output_grads)]
        return StationaryInitGradOp(self.device, self.batched)(T, R, Q, outputs[0], outputs[1], *gb)


class StationaryInitGradOp(Op):
    """T, R, Q, x0, P0, x0_bar, P0_bar -> c_bar, T_bar, R_bar, Q_bar."""

    __props__ = ("device", "batched")

    def __init__(self, device=0, batched=False):
        self.device, self.batched = int(device), bool(batched)

    def make_node(self, T, R, Q, a0, P0, a0_bar, P0_bar):
        ins = [pt.as_tensor_variable(x) for x in (T, R, Q, a0, P0, a0_bar, P0_bar)]
        return Apply(self, ins, [ins[3].type(), ins[0].type(), ins[1].type(), ins[2].type()])

    def perform(self, node, inputs, output_storage, params=None):
        g = default_engine(self.device).stationary_init_grad(*inputs, dtype=node.outputs[0].dtype)
        for store, name in zip(output_storage, ("c", "T", "R", "Q")):
            store[0] = np.asarray(g[name])

    def infer_shape(self, fgraph, node, ins_shapes):
        return [ins_shapes[3], ins_shapes[0], ins_shapes[1], ins_shapes[2]]


# ---- vectorize_graph: draws / series go onto the engine's batch axis ---------------------------------------------
def _flatten_batch(op_batched, node, batched_inputs, what):
    """New ``batched`` flags and inputs for an Op whose inputs gained leading batch dimensions.  One new batch axis per
    input at most, and only on an Op that has none yet (the engine has ONE batch axis: flatten (chain, draw) first)."""
    flags, ins = [], []
    for old, new, bt in zip(node.inputs, batched_inputs, op_batched):
        extra = new.type.ndim - old.type.ndim
        if extra not in (0, 1) or (extra == 1 and bt):
            raise NotImplementedError(
                f"{what}: the engine has one batch axis; reshape the (chain, draw, ...) axes into one before vectorising")
        flags.append(bt or extra == 1)
        ins.append(new)
    return tuple(flags), ins


def _vectorize_filter_op(op, node, *batched_inputs):
    flags, ins = _flatten_batch(op.batched, node, batched_inputs, type(op).__name__)
    if isinstance(op, KalmanGradOp):  # (the ticket and the cotangent ride behind the ten filter inputs)
        flags = flags[:10]
    new_op = type(op)(op.kind, op.cov_jitter, op.missing_fill_value, op.device, op.time_varying, flags)
    return new_op.make_node(*ins)


_vectorize_node.register(KalmanLoglikeOp, _vectorize_filter_op)
_vectorize_node.register(KalmanFilterOp, _vectorize_filter_op)


@_vectorize_node.register(KalmanSmootherOp)
def _vectorize_smoother_op(op, node, *batched_inputs):
    flags, ins = _flatten_batch(op.batched, node, batched_inputs, "KalmanSmootherOp")
    return KalmanSmootherOp(op.cov_jitter, op.device, op.time_varying, flags).make_node(*ins)


@_vectorize_node.register(StationaryInitOp)
def _vectorize_stationary_op(op, node, *batched_inputs):
    extra = [new.type.ndim - old.type.ndim for old, new in zip(node.inputs, batched_inputs)]
    if op.batched or any(e not in (0, 1) for e in extra):
        raise NotImplementedError("StationaryInitOp: the engine has one batch axis")
    if not any(extra):
        return op.make_node(*batched_inputs)
    B = None
    for new, e in zip(batched_inputs, extra):
        if e:
            B = new.shape[0]
    ins = [new if e else pt.broadcast_to(new, (B, *new.shape)) for new, e in zip(batched_inputs, extra)]
    return StationaryInitOp(op.device, True).make_node(*ins)


# ---- graph builders used by filters.py ------------------------------------------------------------------------------
def build_filter_graph(flt, data, a0, P0, c, d, T, Z, R, H, Q):
    """Symbolic branch of ``BaseFilter.build_graph``: outputs named / ordered as the reference does
    (kalman_filter.py:271-308).  The six state outputs and ``loglike_obs`` come from two nodes (module docstring)."""
    data, a0, P0, c, d, T, Z, R, H, Q = (pt.as_tensor_variable(x) for x in (data, a0, P0, c, d, T, Z, R, H, Q))
    R_shape, Z_shape = R.type.shape, Z.type.shape
    flt.n_states, flt.n_shocks = R_shape[-2:]
    flt.n_posdef = flt.n_shocks
    flt.n_endog = Z_shape[-2]
    names = ["c", "d", "T", "Z", "R", "H", "Q"]
    nd = dict(zip(names, (c.ndim, d.ndim, T.ndim, Z.ndim, R.ndim, H.ndim, Q.ndim)))
    flt.seq_names = [k for k in names if nd[k] == _ND[k] + 1]  # decide_if_x_time_varies, filters/utilities.py:8-29
    flt.non_seq_names = [k for k in names if k not in flt.seq_names]
    args = (flt.kind, flt.cov_jitter, flt.missing_fill_value, flt.device, tuple(flt.seq_names))
    states = KalmanFilterOp(*args)(data, a0, P0, c, d, T, Z, R, H, Q)
    ll, _ = KalmanLoglikeOp(*args)(data, a0, P0, c, d, T, Z, R, H, Q)
    outs = [*states, ll]
    for o, name in zip(outs, OUTPUT_NAMES):
        o.name = name
    return outs


def build_smoother_graph(sm, T, R, Q, filtered_states, filtered_covariances):
    T, R, Q = (pt.as_tensor_variable(x) for x in (T, R, Q))
    nd = {"T": T.ndim, "R": R.ndim, "Q": Q.ndim}
    sm.seq_names = [k for k in nd if nd[k] == 3]
    sm.non_seq_names = [k for k in nd if nd[k] != 3]
    sa, sP = KalmanSmootherOp(sm.cov_jitter, sm.device, tuple(sm.seq_names))(
        T, R, Q, filtered_states, filtered_covariances)
    sa.name, sP.name = "smoothed_states", "smoothed_covariances"
    return sa, sP
