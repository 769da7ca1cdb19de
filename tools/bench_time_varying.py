"""Row f-4 in numbers: the structural C3 model with a time-varying design row (RegressionComponent pattern, Z[t],
structural.py:1601-1617) -- batched logp + gradient and filter + smoother on the generic warp-per-series kernels that
stream x[t] per step, beside the same model with a constant Z on the DMMA kernels.  Device-resident inputs.

    python tools/bench_time_varying.py [B] [T]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 240
eng = KalmanEngine(0)
cfg = synth.structural(B=B, n=n)
rng = np.random.default_rng(1)
Zt = np.repeat(cfg["Z"][0][None], n, 0)  # shared over the batch, one row per step
Zt[:, 0, 14] = rng.standard_normal(n)    # an exogenous regressor on the last state
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
base = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
tv = list(base)
tv[1 + PARAM_NAMES.index("Z")] = dev(Zt)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"B": B, "T": n, "m": 15}
for name, args, kw in (("constant_Z", base, {}), ("time_varying_Z", tv, {"time_varying": ("Z",)})):
    ms_g = timeit(lambda: eng.logp_grad("standard", *args, **kw))
    path_g = eng.last_path
    ms_s = timeit(lambda: eng.filter_smooth("standard", *args, **kw))
    out[name] = {"logp_grad_ms": ms_g, "logp_grad_path": path_g, "logp_grad_state_steps_per_s": B * n / ms_g * 1e3,
                 "filter_smooth_ms": ms_s, "filter_smooth_path": eng.last_path,
                 "filter_smooth_state_steps_per_s": B * n / ms_s * 1e3}
print(json.dumps(out))
