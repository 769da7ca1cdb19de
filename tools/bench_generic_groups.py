"""Generic Standard / Univariate kernels (bkf_generic.cu): logp + gradient at every warps-per-series width (BKF_GEN_W = 1, 2,
4, 8) and with the launcher's own choice ("auto"), device-resident inputs.

    python tools/bench_generic_groups.py [quick]
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
eng.enable_timing(True) if hasattr(eng, "enable_timing") else None


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


quick = "quick" in sys.argv
shapes = [(32, 16, 16, 512, 100), (16, 8, 8, 2048, 100), (17, 1, 1, 4096, 100), (27, 1, 1, 4096, 100), (40, 1, 2, 2048, 100),
          (24, 4, 6, 1024, 100)]
widths = ("auto", "1", "8") if quick else ("auto", "1", "2", "4", "8")
rows = []
for (m, p, r, B, n) in shapes:
    cfg = synth.random_dense(B, n, m, p, r, seed=1, missing=0.0)
    args = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
    for kind in ("standard", "univariate"):
        for w in widths:
            if w == "auto":
                os.environ.pop("BKF_GEN_W", None)
            else:
                os.environ["BKF_GEN_W"] = w
            ms = timeit(lambda: eng.logp_grad(kind, *args))
            rows.append({"m": m, "p": p, "r": r, "B": B, "T": n, "filter": kind, "width": w, "path": eng.last_path,
                         "ms": round(ms, 2), "state_steps_per_s": round(B * n / ms * 1e3)})
            print(rows[-1], flush=True)
os.environ.pop("BKF_GEN_W", None)
print(json.dumps(rows))
