"""Two logp + gradient calls on the generic kernels for an ncu capture.

    python tools/dbg_generic_profile.py [kind m p r B n]      (default: standard 32 16 16 296 20, the VARMAX shape)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

a = sys.argv[1:]
kind = a[0] if a else "standard"
m, p, r, B, n = (int(x) for x in a[1:6]) if len(a) >= 6 else (32, 16, 16, 296, 20)
eng = KalmanEngine(0)
cfg = synth.random_dense(B, n, m, p, r, seed=1, missing=0.0)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda:0")
args = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
for _ in range(2):
    eng.logp_grad(kind, *args)
torch.cuda.synchronize()
print(eng.last_path)
