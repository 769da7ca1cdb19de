"""Time bkf_stationary_init / _grad on a C2-sized batch (16 384 models, m = 15, r = 1), device-resident."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pymc_experimental_b200.engine import KalmanEngine  # noqa: E402

eng = KalmanEngine(0)
rng = np.random.default_rng(0)
B, m, r = 16384, 15, 1
T = rng.standard_normal((B, m, m))
T *= (rng.uniform(0.3, 0.98, B) / np.abs(np.linalg.eigvals(T)).max(axis=1))[:, None, None]
dev = lambda a: torch.as_tensor(a, device="cuda:0")
c, T, R, Q = dev(rng.standard_normal((B, m))), dev(T), dev(rng.standard_normal((B, m, r))), dev(np.ones((B, r, r)))
ab, Pb = dev(rng.standard_normal((B, m))), dev(rng.standard_normal((B, m, m)))
for _ in range(3):
    a0, P0, st = eng.stationary_init(c, T, R, Q)
    g = eng.stationary_init_grad(T, R, Q, a0, P0, ab, Pb)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for _ in range(10):
    a0, P0, st = eng.stationary_init(c, T, R, Q)
ev[1].record()
for _ in range(10):
    g = eng.stationary_init_grad(T, R, Q, a0, P0, ab, Pb)
ev[2].record()
torch.cuda.synchronize()
print(f"stationary_init: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms, grad: {ev[1].elapsed_time(ev[2]) / 10:.3f} ms "
      f"for {B} models (m={m}, r={r}); non-stationary flagged: {int((st != 0).sum())}")
