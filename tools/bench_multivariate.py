"""How the filters with p > 1 observed series fare: StandardFilter / UnivariateFilter run on the generic warp-per-series
kernels there, SquareRootFilter on the tile kernels.  Device-resident logp + gradient.

    python tools/bench_multivariate.py
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

eng = KalmanEngine(0)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


def timeit(fn, reps=5):
    """Fastest of reps single calls (the boxes show occasional 1.5x outliers on these long kernels)."""
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


rows = []
for (m, p, r, B, n) in ((32, 16, 16, 512, 200), (16, 8, 8, 2048, 200), (8, 2, 2, 16384, 200), (4, 2, 2, 65536, 100)):
    cfg = synth.random_dense(B, n, m, p, r, seed=1, missing=0.0)
    for kind in ("standard", "univariate", "cholesky"):
        c = dict(cfg)
        if kind == "cholesky":
            c["P0"] = np.linalg.cholesky(cfg["P0"])
        args = [dev(c["y"])] + [dev(c[k]) for k in PARAM_NAMES]
        ms = timeit(lambda: eng.logp_grad(kind, *args))
        rows.append({"m": m, "p": p, "r": r, "B": B, "T": n, "filter": kind, "path": eng.last_path, "ms": round(ms, 2),
                     "state_steps_per_s": round(B * n / ms * 1e3)})
        print(rows[-1], flush=True)
print(json.dumps(rows))
