"""The PyTensor Op layer (pymc_experimental_b200/pytensor_ops.py) driven through the Op protocol:
make_node -> perform -> L_op -> KalmanGradOp.perform, vectorize_graph, pickling.

PyTensor itself is not installable here; when it is missing the protocol stub in tests/stubs/pytensor stands in (its
header says what it implements).  CPU tests run the Ops against an oracle-backed fake engine (the protocol and the
bookkeeping are what they test); the ``gpu`` tests run the same graphs on the real engine against the golden vectors
generated from the reference's own source (tests/golden/generate_golden.py)."""

import os
import pickle
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose

try:  # the real thing when it exists
    import pytensor  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs"))
    import pytensor  # noqa: F401

import pytensor.tensor as pt  # noqa: E402
from pytensor.graph.replace import vectorize_graph  # noqa: E402

from conftest import load_golden  # noqa: E402
from pymc_experimental_b200 import engine as engine_mod  # noqa: E402
from pymc_experimental_b200 import filters, pytensor_ops, stationary  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES  # noqa: E402

ND = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


class FakeEngine:
    """The engine's host API on the CPU oracle: records which calls the Ops make."""

    def __init__(self):
        self.calls = []
        self.saved = None
        self.ticket = 0

    def _batched(self, y, prm):
        B = None
        if np.ndim(y) == 3:
            B = np.shape(y)[0]
        for k, x in zip(PARAM_NAMES, prm):
            if np.ndim(x) == ND[k] + 1:
                B = np.shape(x)[0]
        return B

    def _full(self, y, prm, B):
        yb = np.broadcast_to(y, (B, *np.shape(y)[-2:]))
        return yb, [np.broadcast_to(x, (B, *np.shape(x)[-ND[k]:])) for k, x in zip(PARAM_NAMES, prm)]

    def forward_saved(self, kind, y, *prm, cov_jitter=None, missing_fill_value=None, dtype=None, time_varying=(),
                      want_ll=True):
        from oracle import kalman_np as onp

        self.calls.append("forward_saved")
        B = self._batched(y, prm)
        if B is None:
            ll = onp.kalman_filter(kind, y, *prm, cov_jitter=cov_jitter)[-1]
        else:
            yb, pb = self._full(y, prm, B)
            ll = np.stack([onp.kalman_filter(kind, yb[b], *[x[b] for x in pb], cov_jitter=cov_jitter)[-1] for b in range(B)])
        self.ticket += 1
        self.saved = (self.ticket, kind, y, prm, cov_jitter)
        return ll, ll.sum(-1), np.zeros(B or (), dtype=np.int32), self.ticket

    def _grads(self, kind, y, prm, cot, cov_jitter):
        from oracle import kalman_torch as oth

        B = self._batched(y, prm)
        if B is None:
            _, g = oth.logp_and_grad(kind, y[None], *[np.asarray(x)[None] for x in prm], ll_cotangent=cot[None],
                                     cov_jitter=cov_jitter)
            return {k: v[0] for k, v in g.items()}
        yb, pb = self._full(y, prm, B)
        _, g = oth.logp_and_grad(kind, np.array(yb), *[np.array(x) for x in pb], ll_cotangent=cot, cov_jitter=cov_jitter)
        return g

    def backward_saved(self, ticket, ll_cotangent=None, grads=PARAM_NAMES):
        self.calls.append("backward_saved")
        saved, self.saved = self.saved, None
        if saved is None or saved[0] != ticket:
            return None
        _, kind, y, prm, jit = saved
        return self._grads(kind, y, prm, ll_cotangent, jit)

    def logp_grad(self, kind, y, *prm, ll_cotangent=None, cov_jitter=None, missing_fill_value=None, dtype=None,
                  time_varying=()):
        self.calls.append("logp_grad")
        g = self._grads(kind, y, prm, ll_cotangent, cov_jitter)
        return None, g, None

    def filter(self, kind, y, *prm, outputs=None, cov_jitter=None, missing_fill_value=None, dtype=None, time_varying=()):
        from oracle import kalman_np as onp

        self.calls.append("filter")
        outs = onp.kalman_filter(kind, y, *prm, cov_jitter=cov_jitter)
        return dict(zip(pytensor_ops.OUTPUT_NAMES, outs))

    def smooth(self, T, R, Q, fa, fP, cov_jitter=None, dtype=None, time_varying=()):
        from oracle import kalman_np as onp

        self.calls.append("smooth")
        return onp.kalman_smoother(T, R, Q, fa, fP, cov_jitter=cov_jitter)

    def stationary_init(self, c, T, R, Q, dtype=None):
        from oracle import stationary as ost

        self.calls.append("stationary_init")
        if np.ndim(T) == 3:
            res = [ost.stationary_init_np(c[b], T[b], R[b], Q[b]) for b in range(len(T))]
            return np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.zeros(len(T), dtype=np.int32)
        a0, P0 = ost.stationary_init_np(c, T, R, Q)
        return a0, P0, np.int32(0)

    def stationary_init_grad(self, T, R, Q, a0, P0, a0_bar, P0_bar, dtype=None):
        from oracle import stationary as ost

        self.calls.append("stationary_init_grad")
        c = (np.eye(len(T)) - T) @ a0
        return ost.stationary_init_grad(c, T, R, Q, a0_bar, P0_bar)[2]


@pytest.fixture
def fake(monkeypatch):
    eng = FakeEngine()
    monkeypatch.setattr(pytensor_ops, "default_engine", lambda device=0: eng)
    return eng


def symbols(ins, batched=()):
    """One symbolic input per golden input, static shapes as the reference's graphs have them."""
    sv = {"y": pt.tensor("data", shape=(None, ins["y"].shape[-1]))}
    for k in PARAM_NAMES:
        sv[k] = pt.tensor(k, shape=ins[k].shape)
    return sv


def build(kind, sv):
    flt = filters.FILTER_FACTORY[kind]()
    outs = flt.build_graph(sv["y"], *[sv[k] for k in PARAM_NAMES])
    return flt, outs


def values(ins):
    return [ins["y"]] + [ins[k] for k in PARAM_NAMES]


# ---- CPU: protocol -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["dense_m5p2r3_standard", "c2_sarima_univariate", "dense_m2p1r2_cholesky"])
def test_graph_outputs_and_attributes_match_the_reference_build_graph(fake, name):
    kind, ins, ref = load_golden(name)
    sv = symbols(ins)
    flt, outs = build(kind, sv)
    assert [o.name for o in outs] == pytensor_ops.OUTPUT_NAMES  # kalman_filter.py:298-308
    m, r = ins["R"].shape
    assert (flt.n_states, flt.n_shocks, flt.n_endog) == (m, r, ins["Z"].shape[0])
    assert flt.seq_names == [] and flt.non_seq_names == ["c", "d", "T", "Z", "R", "H", "Q"]
    assert flt.cov_jitter == 1e-8  # floatX float64 (utils/constants.py:20)
    assert outs[0].type.shape == (None, m) and outs[3].type.shape == (None, m, m) and outs[6].type.shape == (None,)
    f = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], outs)
    got = f(*values(ins))
    for o, n in zip(got, pytensor_ops.OUTPUT_NAMES):
        assert_allclose(o, ref[n], rtol=1e-8, atol=1e-8 * np.abs(ref[n]).max(), err_msg=n)


def test_logp_only_function_runs_the_forward_sweep_once_and_nothing_else(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    f = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], outs[6].sum())
    assert_allclose(f(*values(ins)), ref["loglike_obs"].sum(), rtol=1e-10)
    assert fake.calls == ["forward_saved"]  # the six-output filter node was pruned: no state / covariance copies


def test_states_only_function_does_not_run_the_loglike_node(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    f = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], [outs[0], outs[3]])
    f(*values(ins))
    assert fake.calls == ["filter"]


@pytest.mark.parametrize("name", ["dense_m5p2r3_standard", "c2_sarima_univariate", "dense_m2p1r2_cholesky"])
def test_grad_through_the_op_one_forward_one_adjoint(fake, name):
    kind, ins, ref = load_golden(name)
    sv = symbols(ins)
    _, outs = build(kind, sv)
    logp = outs[6].sum()
    wrt = [sv[k] for k in PARAM_NAMES]
    grads = pytensor.grad(logp, wrt)  # the connection_pattern / DisconnectedType consistency check runs here
    f = pytensor.function([sv["y"]] + wrt, [logp, *grads])
    got = f(*values(ins))
    assert_allclose(got[0], ref["loglike_obs"].sum(), rtol=1e-10)
    for k, g in zip(PARAM_NAMES, got[1:]):
        assert g.shape == ins[k].shape
        assert_allclose(g, ref["grad_" + k], rtol=0, atol=1e-8 * max(np.abs(ref["grad_" + k]).max(), 1e-300), err_msg=k)
    assert fake.calls == ["forward_saved", "backward_saved"]  # ONE forward sweep per logp + dlogp evaluation


def test_weighted_cotangent_reaches_the_adjoint(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    w = np.linspace(0.5, 1.5, ins["y"].shape[0])
    cost = (outs[6] * w).sum()
    g_T = pytensor.grad(cost, sv["T"])
    got = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], g_T)(*values(ins))
    from oracle import kalman_torch as oth

    _, g = oth.logp_and_grad(kind, ins["y"][None], *[ins[k][None] for k in PARAM_NAMES], ll_cotangent=w[None])
    assert_allclose(got, g["T"][0], rtol=0, atol=1e-9 * np.abs(g["T"]).max())


def test_the_data_is_a_disconnected_input_not_an_undefined_one(fake):
    kind, ins, _ = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    node = outs[6].owner
    assert node.op.connection_pattern(node)[0] == [False, False]
    from pytensor.gradient import DisconnectedType

    igs = node.op.L_op(node.inputs, node.outputs, [pt.ones_like(outs[6]), pytensor.gradient.disconnected_type()])
    assert isinstance(igs[0].type, DisconnectedType)
    assert all(not isinstance(g.type, DisconnectedType) for g in igs[1:])
    with pytest.raises(Exception):  # asking for d/d data is an error, as for any disconnected input
        pytensor.grad(outs[6].sum(), sv["y"])
    # gradient through a state output is refused loudly
    with pytest.raises(NotImplementedError):
        pytensor.grad(outs[0].sum(), sv["T"])


def test_stale_ticket_falls_back_to_the_fused_call(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    wrt = [sv[k] for k in PARAM_NAMES]
    grads = pytensor.grad(outs[6].sum(), wrt)
    f = pytensor.function([sv["y"]] + wrt, grads)
    orig = fake.forward_saved

    def forward_then_lose_the_record(*a, **k):
        res = orig(*a, **k)
        fake.saved = None  # something else used the engine's workspace
        return res

    fake.forward_saved = forward_then_lose_the_record
    got = f(*values(ins))
    assert fake.calls == ["forward_saved", "backward_saved", "logp_grad"]
    assert_allclose(got[4], ref["grad_T"], rtol=0, atol=1e-8 * np.abs(ref["grad_T"]).max())


def test_time_varying_inputs_are_detected_by_rank_as_the_reference_does(fake):
    kind, ins, _ = load_golden("dense_m5p2r3_standard")
    n = ins["y"].shape[0]
    sv = symbols(ins)
    sv["Z"] = pt.tensor("Z", shape=(n, *ins["Z"].shape))
    flt, outs = build(kind, sv)
    assert flt.seq_names == ["Z"] and "Z" not in flt.non_seq_names  # filters/utilities.py:8-29
    assert outs[6].owner.op.time_varying == ("Z",)


def test_vectorize_graph_maps_draws_onto_the_batch_axis(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    B = 3
    rng = np.random.default_rng(0)
    Tb = pt.tensor("T_draws", shape=(B, *ins["T"].shape))
    Qb = pt.tensor("Q_draws", shape=(B, *ins["Q"].shape))
    ll_b = vectorize_graph(outs[6], {sv["T"]: Tb, sv["Q"]: Qb})
    op = ll_b.owner.op
    assert isinstance(op, pytensor_ops.KalmanLoglikeOp) and op.batch_ndim == 1
    assert op.batched == (False, False, False, False, False, True, False, False, False, True)
    assert ll_b.type.shape == (B, None)
    Tv = ins["T"][None] * (1 + 0.01 * rng.standard_normal((B, 1, 1)))
    Qv = ins["Q"][None] * (1 + 0.1 * rng.uniform(size=(B, 1, 1)))
    others = [sv["y"]] + [sv[k] for k in PARAM_NAMES if k not in ("T", "Q")]
    f = pytensor.function(others + [Tb, Qb], ll_b)
    got = f(*[ins["y"]] + [ins[k] for k in PARAM_NAMES if k not in ("T", "Q")] + [Tv, Qv])
    from oracle import kalman_np as onp

    for b in range(B):
        p = dict(ins, T=Tv[b], Q=Qv[b])
        assert_allclose(got[b], onp.kalman_filter(kind, p["y"], *[p[k] for k in PARAM_NAMES])[-1], rtol=1e-10)
    assert fake.calls == ["forward_saved"]  # one batched call, not a loop over draws
    # gradients of a batched node: per-draw for batched inputs, summed over the draws for shared ones
    gT, gR = pytensor.grad(ll_b.sum(), [Tb, sv["R"]])
    assert gT.type.shape == Tb.type.shape and gR.type.shape == sv["R"].type.shape
    gTv, gRv = pytensor.function(others + [Tb, Qb], [gT, gR])(
        *[ins["y"]] + [ins[k] for k in PARAM_NAMES if k not in ("T", "Q")] + [Tv, Qv])
    from oracle import kalman_torch as oth

    full = [np.broadcast_to(ins[k], (B, *ins[k].shape)).copy() for k in PARAM_NAMES]
    full[PARAM_NAMES.index("T")], full[PARAM_NAMES.index("Q")] = Tv, Qv
    _, g = oth.logp_and_grad(kind, np.broadcast_to(ins["y"], (B, *ins["y"].shape)).copy(), *full)
    assert_allclose(gTv, g["T"], rtol=0, atol=1e-9 * np.abs(g["T"]).max())
    assert_allclose(gRv, g["R"].sum(0), rtol=0, atol=1e-9 * np.abs(g["R"]).max())
    # a second batch axis has to be flattened by the caller
    with pytest.raises(NotImplementedError):
        vectorize_graph(ll_b, {Tb: pt.tensor("T2", shape=(2, B, *ins["T"].shape))})


def test_smoother_and_stationary_ops(fake):
    kind, ins, ref = load_golden("dense_m5p2r3_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    sm = filters.KalmanSmoother()
    sa, sP = sm.build_graph(sv["T"], sv["R"], sv["Q"], outs[0], outs[3])
    assert (sa.name, sP.name) == ("smoothed_states", "smoothed_covariances")
    got = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], [sa, sP])(*values(ins))
    assert_allclose(got[0], ref["smoothed_states"], rtol=0, atol=1e-8 * np.abs(ref["smoothed_states"]).max())
    assert_allclose(got[1], ref["smoothed_covariances"], rtol=0, atol=1e-8 * np.abs(ref["smoothed_covariances"]).max())
    # stationary initialisation + its pull-back (models/SARIMAX.py:369-381)
    rng = np.random.default_rng(1)
    m, r = 4, 2
    T = 0.5 * rng.standard_normal((m, m)) / np.sqrt(m)
    R, Qh = rng.standard_normal((m, r)), rng.standard_normal((r, r))
    Q, c = Qh @ Qh.T + np.eye(r), rng.standard_normal(m)
    cs, Ts, Rs, Qs = pt.tensor("c", shape=(m,)), pt.tensor("T", shape=(m, m)), pt.tensor("R", shape=(m, r)), pt.tensor("Q", shape=(r, r))
    x0, P0 = stationary.stationary_initialization(cs, Ts, Rs, Qs)
    cost = x0.sum() + (P0 * P0).sum()
    gs = pytensor.grad(cost, [cs, Ts, Rs, Qs])
    vals = pytensor.function([cs, Ts, Rs, Qs], [x0, P0, *gs])(c, T, R, Q)
    from oracle import stationary as ost

    a_ref, P_ref = ost.stationary_init_np(c, T, R, Q)
    assert_allclose(vals[0], a_ref, rtol=1e-9)
    assert_allclose(vals[1], P_ref, rtol=1e-9)
    g_ref = ost.stationary_init_grad(c, T, R, Q, np.ones(m), 2 * P_ref)[2]
    for got_g, k in zip(vals[2:], ("c", "T", "R", "Q")):
        assert_allclose(got_g, g_ref[k], rtol=0, atol=1e-8 * np.abs(g_ref[k]).max())


def test_ops_are_picklable_and_compare_by_props():
    a = pytensor_ops.KalmanLoglikeOp("univariate", 1e-8, -9999.0, 0, ("Z",))
    b = pickle.loads(pickle.dumps(a))
    assert a == b and hash(a) == hash(b) and b.time_varying == ("Z",) and b.batched == (False,) * 10
    assert a != pytensor_ops.KalmanLoglikeOp("standard", 1e-8, -9999.0, 0, ("Z",))
    for op in (pytensor_ops.KalmanFilterOp("standard", 1e-8, -9999.0), pytensor_ops.KalmanGradOp("standard", 1e-8, -9999.0),
               pytensor_ops.KalmanSmootherOp(1e-8), pytensor_ops.StationaryInitOp(), pytensor_ops.StationaryInitGradOp()):
        assert pickle.loads(pickle.dumps(op)) == op


def test_float32_graphs_get_the_float32_jitter(fake, monkeypatch):
    """utils/constants.py:20: JITTER_DEFAULT = 1e-8 if floatX == float64 else 1e-6."""
    monkeypatch.setattr(pytensor.config, "floatX", "float32")
    kind, ins, _ = load_golden("dense_m5p2r3_standard")
    sv = {"y": pt.tensor("data", dtype="float32", shape=(None, ins["y"].shape[-1]))}
    for k in PARAM_NAMES:
        sv[k] = pt.tensor(k, dtype="float32", shape=ins[k].shape)
    flt, outs = build(kind, sv)
    assert flt.cov_jitter == 1e-6 and outs[6].owner.op.cov_jitter == 1e-6 and outs[6].type.dtype == "float32"
    flt2, outs2 = filters.StandardFilter(), None
    outs2 = flt2.build_graph(sv["y"], *[sv[k] for k in PARAM_NAMES], cov_jitter=3e-7)  # an explicit value is kept
    assert flt2.cov_jitter == 3e-7 and outs2[6].owner.op.cov_jitter == 3e-7


def test_engine_refuses_to_run_in_a_child_forked_after_cuda_was_initialised(monkeypatch):
    monkeypatch.setattr(engine_mod, "_cuda_pid", os.getpid() + 12345)  # as if our parent had created the engine
    monkeypatch.setattr(engine_mod, "_default_engines", {})
    with pytest.raises(engine_mod.BkfError, match="spawn"):
        engine_mod.default_engine(0)


def test_fork_drops_inherited_handles(monkeypatch):
    class H:
        h = object()

    eng = H()
    monkeypatch.setattr(engine_mod, "_default_engines", {0: eng})
    monkeypatch.setattr(engine_mod, "_pinned_pool", object())
    engine_mod._after_fork_in_child()
    assert eng.h is None and engine_mod._default_engines == {} and engine_mod._pinned_pool is None


# ---- GPU: the same graphs on the real engine ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def real_engine():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return engine_mod.default_engine(0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_sarima_univariate", "c3_structural_standard", "c4_varmax_k4_cholesky",
                                  "dense_m5p2r3_standard", "dense_m15p1r5_missing_univariate"])
def test_gpu_logp_and_grad_through_the_ops_two_launches(real_engine, name):
    kind, ins, ref = load_golden(name)
    sv = symbols(ins)
    _, outs = build(kind, sv)
    wrt = [sv[k] for k in PARAM_NAMES]
    logp = outs[6].sum()
    f = pytensor.function([sv["y"]] + wrt, [logp, *pytensor.grad(logp, wrt)])
    f(*values(ins))
    n0 = real_engine.launch_count
    got = f(*values(ins))
    assert real_engine.launch_count - n0 == 2  # one forward kernel, one adjoint kernel per logp + dlogp evaluation
    assert_allclose(got[0], ref["loglike_obs"].sum(), rtol=1e-9)
    for k, g in zip(PARAM_NAMES, got[1:]):
        assert_allclose(g, ref["grad_" + k], rtol=0, atol=1e-8 * max(np.abs(ref["grad_" + k]).max(), 1e-300), err_msg=k)


@pytest.mark.gpu
def test_gpu_all_ops_against_the_golden_outputs(real_engine):
    kind, ins, ref = load_golden("c3_structural_standard")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    sa, sP = filters.KalmanSmoother().build_graph(sv["T"], sv["R"], sv["Q"], outs[0], outs[3])
    got = pytensor.function([sv["y"]] + [sv[k] for k in PARAM_NAMES], [*outs, sa, sP])(*values(ins))
    for o, n in zip(got, pytensor_ops.OUTPUT_NAMES + ["smoothed_states", "smoothed_covariances"]):
        assert_allclose(o, ref[n], rtol=0, atol=1e-8 * np.abs(ref[n]).max(), err_msg=n)


@pytest.mark.gpu
def test_gpu_saved_forward_then_adjoint_equals_the_fused_call_and_goes_stale(real_engine):
    kind, ins, ref = load_golden("c2_sarima_univariate")
    args = values(ins)
    ll, logp, status, ticket = real_engine.forward_saved(kind, *args)
    g = real_engine.backward_saved(ticket)
    logp2, g2, _ = real_engine.logp_grad(kind, *args)
    assert logp == logp2 and all(np.array_equal(g[k], g2[k]) for k in PARAM_NAMES)  # bit-identical to the fused call
    assert real_engine.backward_saved(ticket) is None  # one ticket, one adjoint
    _, _, _, ticket = real_engine.forward_saved(kind, *args)
    real_engine.filter(kind, *args)  # anything that uses the workspace invalidates the record ...
    assert real_engine.backward_saved(ticket) is None  # ... and the adjoint refuses instead of reading garbage


@pytest.mark.gpu
def test_gpu_vectorized_graph_equals_a_loop_over_draws(real_engine):
    kind, ins, _ = load_golden("c2_sarima_univariate")
    sv = symbols(ins)
    _, outs = build(kind, sv)
    B = 7
    rng = np.random.default_rng(3)
    Qb = pt.tensor("Q_draws", shape=(B, *ins["Q"].shape))
    ll_b = vectorize_graph(outs[6], {sv["Q"]: Qb})
    gQ = pytensor.grad(ll_b.sum(), Qb)
    others = [sv["y"]] + [sv[k] for k in PARAM_NAMES if k != "Q"]
    Qv = ins["Q"][None] * rng.uniform(0.5, 2.0, size=(B, 1, 1))
    f = pytensor.function(others + [Qb], [ll_b, gQ])
    n0 = real_engine.launch_count
    llv, gv = f(*[ins["y"]] + [ins[k] for k in PARAM_NAMES if k != "Q"] + [Qv])
    assert real_engine.launch_count - n0 == 2
    for b in range(B):
        p = dict(ins, Q=Qv[b])
        logp, g, _ = real_engine.logp_grad(kind, p["y"], *[p[k] for k in PARAM_NAMES])
        assert_allclose(llv[b].sum(), logp, rtol=1e-12)
        assert_allclose(gv[b], g["Q"], rtol=1e-10)


@pytest.mark.gpu
def test_gpu_eager_float32_inputs_get_the_float32_jitter(real_engine):
    import torch

    kind, ins, _ = load_golden("dense_m2p1r2_standard")
    t32 = [torch.as_tensor(x, dtype=torch.float32, device="cuda") for x in values(ins)]
    flt = filters.StandardFilter()
    flt.build_graph(*t32)
    assert flt.cov_jitter == 1e-6
    flt.build_graph(*values(ins))
    assert flt.cov_jitter == 1e-8
    with pytest.raises(engine_mod.BkfError):  # tensors of another device are refused, not run on the wrong GPU
        class Dev:
            index = 5
        bufs = type("B", (), {"dev": Dev()})()
        real_engine.use_torch_stream(bufs)
