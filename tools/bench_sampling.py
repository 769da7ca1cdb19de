"""Timing of bkf_mvn_sample at the C3 size: one draw per (draw, step) from the smoother's
output, device-resident inputs, CUDA events around the launch.   python tools/bench_sampling.py [B] [T]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymc_experimental_b200 import synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 240
eng = KalmanEngine(0)
cfg = synth.structural(B=B, n=T)
d = {k: torch.as_tensor(v, dtype=torch.float64, device="cuda:0") for k, v in cfg.items()}
ll, sa, sP, st = eng.filter_smooth("standard", d["y"], *[d[k] for k in PARAM_NAMES])
z = torch.randn(sa.shape, dtype=torch.float64, device="cuda:0")
for _ in range(3):
    x, _ = eng.sample_mvn(sa, sP, z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    x, _ = eng.sample_mvn(sa, sP, z)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
m = sa.shape[-1]
bytes_ = B * T * (m * m + 3 * m) * 8  # cov + mu + z read, x written
print(json.dumps({"workload": f"C3 smoother output, {B} x {T}, m={m}: one MvNormal draw per (draw, step)", "ms": ms,
                  "draws_per_s": B * T / ms * 1e3, "algorithmic_bytes": bytes_, "achieved_GBs": bytes_ / ms / 1e6,
                  "hbm_peak_GBs": 6538.9, "frac": bytes_ / ms / 1e6 / 6538.9, "path": eng.last_path}))

# unconditional simulation (bkf_simulate) of the same models: B series x T steps
g = torch.Generator(device="cuda:0").manual_seed(3)
r, p = d["R"].shape[-1], d["Z"].shape[-2]
zs = [torch.randn(sh, dtype=torch.float64, device="cuda:0", generator=g) for sh in ((B, m), (B, p), (B, T, r), (B, T, p))]
for _ in range(3):
    xs, ys, _ = eng.simulate(*[d[k] for k in PARAM_NAMES], *zs)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    xs, ys, _ = eng.simulate(*[d[k] for k in PARAM_NAMES], *zs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"workload": f"C3 model, {B} series x {T} steps: unconditional simulation (states + observations)",
                  "ms": ms, "state_steps_per_s": B * T / ms * 1e3, "path": eng.last_path}))
