"""Generic smoother kernel (bkf_generic.cu smooth_series) at state widths without a specialised smoother: device time of
the smoother launch with the Cholesky solve (default) and with the eigen-decomposition pinv at every step
(BKF_GEN_SMOOTH_PINV=1, what the kernel did before), and the largest difference between the two results.

    python tools/bench_generic_smoother.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

eng = KalmanEngine(0)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
rows = []
for (m, p, r, B, n) in ((17, 1, 1, 1024, 50), (27, 1, 1, 1024, 50), (32, 16, 16, 512, 50), (40, 1, 2, 512, 50)):
    cfg = synth.random_dense(B, n, m, p, r, seed=1, missing=0.0)
    args = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
    out = eng.filter("univariate", *args, outputs=("filtered_states", "filtered_covariances"))
    fa, fP = out["filtered_states"], out["filtered_covariances"]
    res = {}
    for mode in ("chol", "pinv"):
        if mode == "pinv":
            os.environ["BKF_GEN_SMOOTH_PINV"] = "1"
        else:
            os.environ.pop("BKF_GEN_SMOOTH_PINV", None)
        eng.smooth(args[5], args[7], args[9], fa, fP)
        eng.enable_timing(True)
        sa, sP = eng.smooth(args[5], args[7], args[9], fa, fP)
        torch.cuda.synchronize()
        ms = eng.kernel_times()["smoother"][0]
        eng.enable_timing(False)
        res[mode] = (ms, torch.as_tensor(sa).cpu().numpy(), torch.as_tensor(sP).cpu().numpy())
    os.environ.pop("BKF_GEN_SMOOTH_PINV", None)
    da = np.abs(res["chol"][1] - res["pinv"][1]).max() / np.abs(res["pinv"][1]).max()
    dP = np.abs(res["chol"][2] - res["pinv"][2]).max() / np.abs(res["pinv"][2]).max()
    rows.append({"m": m, "B": B, "T": n, "ms_chol": round(res["chol"][0], 2), "ms_pinv": round(res["pinv"][0], 2),
                 "state_steps_per_s_chol": round(B * n / res["chol"][0] * 1e3), "max_rel_diff_states": float(da),
                 "max_rel_diff_covariances": float(dP)})
    print(rows[-1], flush=True)
print(json.dumps(rows))
