"""Ill-conditioned P_hat (one shock, m = 27: the small eigenvalues of P_hat are the 1e-8 jitter): errors of the generic
smoother's two paths against the numpy oracle (SVD pinv) and against each other, with cond(P_hat) of the worst step."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import kalman_np as onp  # noqa: E402
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

eng = KalmanEngine(0)
for (m, r) in ((27, 1), (27, 27), (40, 2)):
    cfg = synth.random_dense(3, 20, m, 1, r, seed=1, missing=0.0)
    args = [cfg["y"]] + [cfg[k] for k in PARAM_NAMES]
    out = eng.filter("univariate", *args, outputs=("filtered_states", "filtered_covariances"))
    fa, fP = out["filtered_states"], out["filtered_covariances"]
    res = {}
    for mode in ("chol", "pinv"):
        os.environ["BKF_GEN_SMOOTH_PINV"] = "1" if mode == "pinv" else "0"
        res[mode] = eng.smooth(cfg["T"], cfg["R"], cfg["Q"], fa, fP)
    ra, rP = zip(*[onp.kalman_smoother(cfg["T"][b], cfg["R"][b], cfg["Q"][b], fa[b], fP[b]) for b in range(3)])
    ra, rP = np.stack(ra), np.stack(rP)
    T, R, Q = cfg["T"][0], cfg["R"][0], cfg["Q"][0]
    conds = [np.linalg.cond(T @ fP[0][t] @ T.T + R @ Q @ R.T + 1e-8 * np.eye(m)) for t in range(19)]
    rel = lambda x, y: float(np.abs(x - y).max() / np.abs(y).max())
    print({"m": m, "r": r, "cond_P_hat_max": f"{max(conds):.1e}",
           "chol_vs_oracle": (f"{rel(res['chol'][0], ra):.1e}", f"{rel(res['chol'][1], rP):.1e}"),
           "pinv_vs_oracle": (f"{rel(res['pinv'][0], ra):.1e}", f"{rel(res['pinv'][1], rP):.1e}"),
           "chol_vs_pinv": (f"{rel(res['chol'][0], res['pinv'][0]):.1e}", f"{rel(res['chol'][1], res['pinv'][1]):.1e}")})
