"""Generate tests/golden/*.npz by executing the REFERENCE's own filter source through the eager PyTensor shim.

Run in the build container only (needs /root/reference):

    python tests/golden/generate_golden.py

For every case the script stores the inputs, the seven outputs of ``<Filter>().build_graph`` (reference code:
pymc_extras/statespace/filters/kalman_filter.py), the two outputs of ``KalmanSmoother().build_graph``
(kalman_smoother.py) and, via torch.autograd through that same reference code, the gradients of
``loglike_obs.sum()`` with respect to the nine state-space inputs (what ``pytensor.grad`` yields for NUTS).

Cases:
  nile_*      the reference's statsmodels test inputs (tests/statespace/test_kalman_filter.py:232-269,
              utilities/test_helpers.py:138-170, test_data/nile.csv), n_missing in {0, 5}
  shapes_*    make_test_inputs shapes (utilities/test_helpers.py:86-111), incl. p=5 and missing rows
  dense_*     seeded dense random systems (all nine inputs dense, c,d != 0), with/without missing values
  c2/c3/c4/c5 reduced-size instances of the BASELINE.json configs (pymc_experimental_b200/synth.py)
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ptshim  # noqa: E402
from pymc_experimental_b200 import synth  # noqa: E402

REF = "/root/reference"
kf, ks = ptshim.load_reference_filters(REF)
FILTERS = {"standard": kf.StandardFilter, "univariate": kf.UnivariateFilter, "cholesky": kf.SquareRootFilter}
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
OUT_NAMES = ("filtered_states", "predicted_states", "observed_states", "filtered_covariances",
             "predicted_covariances", "observed_covariances", "loglike_obs")


def run_reference(kind, y, params, smoother=True, grad=True):
    ins = [ptshim.to_var(y)] + [ptshim.to_var(params[k], requires_grad=grad) for k in NAMES]
    outs = FILTERS[kind]().build_graph(*ins)
    res = {n: o.v.detach().numpy().copy() for n, o in zip(OUT_NAMES, outs)}
    if smoother:
        sm = ks.KalmanSmoother().build_graph(ins[5], ins[7], ins[9], outs[0], outs[3])
        res["smoothed_states"] = sm[0].v.detach().numpy().copy()
        res["smoothed_covariances"] = sm[1].v.detach().numpy().copy()
    if grad:
        gs = torch.autograd.grad(outs[6].v.sum(), [v.v for v in ins[1:]], allow_unused=True)
        for k, g, v in zip(NAMES, gs, ins[1:]):
            res["grad_" + k] = (torch.zeros_like(v.v) if g is None else g).numpy().copy()
    return res


def save(name, kind, y, params, extra=None, **kw):
    res = run_reference(kind, y, params, **kw)
    payload = {"in_y": y, "kind": np.array(kind)}
    payload.update(extra or {})
    payload.update({"in_" + k: np.asarray(params[k], dtype=np.float64) for k in NAMES})
    payload.update({"out_" + k: v for k, v in res.items()})
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **payload)
    print(f"{name:34s} logp = {res['loglike_obs'].sum(): .12f}")


def nile_inputs(n_missing, rng):
    import pandas as pd

    x = pd.read_csv(f"{REF}/tests/statespace/test_data/nile.csv")["x"].values.astype(np.float64)
    y = ((x - x.mean()) / x.std(ddof=1))[:, None]  # test_helpers.py:32 (pandas std, ddof=1)
    if n_missing:
        y[rng.choice(len(y), n_missing, replace=False)] = np.nan
    params = dict(a0=np.zeros(2), P0=np.eye(2) * 1e6, c=np.zeros(2), d=np.zeros(1),
                  T=np.array([[1.0, 1.0], [0.0, 1.0]]), Z=np.array([[1.0, 0.0]]), R=np.eye(2),
                  H=np.eye(1) * 0.8, Q=np.diag([0.5, 0.01]))
    return y, params


def shapes_inputs(p, m, r, n, rng, missing=None, H_is_zero=False):
    """utilities/test_helpers.py:86-111 (make_test_inputs)."""
    y = np.arange(n * p, dtype=np.float64).reshape(-1, p)
    if missing:
        y[rng.choice(n, missing, replace=False)] = np.nan
    T = np.eye(m, k=-1)
    T[0, :] = 1 / m
    params = dict(a0=np.zeros(m), P0=np.eye(m), c=np.zeros(m), d=np.zeros(p), T=T, Z=np.eye(m)[:p, :],
                  R=np.eye(m)[:, :r], H=np.zeros((p, p)) if H_is_zero else np.eye(p), Q=np.eye(r))
    return y, params


def one(cfg, i=0):
    """Series i of a batched synth config -> (y, params)."""
    y = cfg["y"] if cfg["y"].ndim == 2 else cfg["y"][i]
    return y, {k: cfg[k][i] for k in NAMES}


TV_CASES = [
    # name, (m, p, r, n), time-varying inputs, missing fraction
    ("tv_all_m4p2r2", (4, 2, 2, 18), ("c", "d", "T", "Z", "R", "H", "Q"), 0.0),
    ("tv_all_m4p2r2_missing", (4, 2, 2, 18), ("c", "d", "T", "Z", "R", "H", "Q"), 0.15),
    ("tv_regressionZ_m5p1r2", (5, 1, 2, 24), ("Z",), 0.0),        # structural.py:1601-1617 (RegressionComponent)
    ("tv_TQ_m3p1r1", (3, 1, 1, 20), ("T", "Q"), 0.1),
]


def tv_cases():
    """Time-varying inputs (leading time axis, filters/utilities.py:8-47), StandardFilter only: the reference's
    UnivariateFilter.kalman_step (kalman_filter.py:791) takes its arguments positionally without unpack_args, so with
    any time-varying input the scan hands it (y, *seqs, a, P, *non_seqs) in the wrong slots and it fails -- there is
    no reference behaviour to record for that combination."""
    for name, (m, p, r, n), tv, miss in TV_CASES:
        cfg = synth.random_dense_tv(1, n, m, p, r, tv, seed=31 * m + p + len(tv), missing=miss, diag_H=bool(miss) and p > 1)
        y, prm = one(cfg)
        save(f"{name}_standard", "standard", y, prm, extra={"time_varying": np.array(tv)})


def main():
    if "--only-tv" in sys.argv:
        return tv_cases()
    rng = np.random.default_rng(sum(map(ord, "statespace")))  # tests/statespace/utilities/shared_fixtures.py:4
    for n_missing in (0, 5):
        y, prm = nile_inputs(n_missing, rng)
        for kind in FILTERS:
            q = dict(prm)
            if kind == "cholesky":
                q["P0"] = np.linalg.cholesky(prm["P0"])  # test_kalman_filter.py:234-235
            save(f"nile_{kind}_miss{n_missing}", kind, y, q)

    for kind in FILTERS:
        y, prm = shapes_inputs(1, 5, 1, 10, rng)
        save(f"shapes_p1m5_{kind}", kind, y, prm)
        y, prm = shapes_inputs(5, 5, 1, 10, rng, missing=1)
        save(f"shapes_p5m5_missing_{kind}", kind, y, prm)

    for (m, p, r, n) in [(2, 1, 2, 30), (5, 2, 3, 25), (15, 1, 5, 40), (6, 3, 2, 20)]:
        for miss in (0.0, 0.15):
            cfg = synth.random_dense(1, n, m, p, r, seed=100 * m + p, missing=miss, diag_H=bool(miss) and p > 1)
            y, prm = one(cfg)
            tag = f"dense_m{m}p{p}r{r}" + ("_missing" if miss else "")
            for kind in FILTERS:
                if kind == "cholesky" and miss and p > 1:
                    # reference fails here: cholesky(W H) of a row-masked H is not PD (kalman_filter.py:666)
                    continue
                save(f"{tag}_{kind}", kind, y, prm)

    y, prm = one(synth.sarima(B=2, n=60), 1)
    save("c2_sarima_univariate", "univariate", y, prm)
    save("c2_sarima_standard", "standard", y, prm)
    y, prm = one(synth.structural(B=2, n=48), 1)
    save("c3_structural_standard", "standard", y, prm)
    y, prm = one(synth.varmax(B=2, n=30, k=4), 1)
    save("c4_varmax_k4_cholesky", "cholesky", y, prm)
    y, prm = one(synth.ets(B=2, n=50), 1)
    save("c5_ets_standard", "standard", y, prm)
    y, prm = one(synth.local_level(B=1, n=100), 0)
    save("c1_local_level_standard", "standard", y, prm)
    tv_cases()


if __name__ == "__main__":
    main()
