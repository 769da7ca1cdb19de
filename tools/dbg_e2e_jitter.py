"""Per-step wall times of the host-buffer logp+grad call (C2 shape): where does the run-to-run spread of `e2e` come from?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymc_experimental_b200 import synth
from pymc_experimental_b200.engine import KalmanEngine, PARAM_NAMES
eng = KalmanEngine(0)
cfg = synth.sarima(B=16384, n=500)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.copy_(torch.as_tensor(a)); return t.numpy()
hargs = [pinned(cfg["y"])] + [pinned(cfg[k]) for k in PARAM_NAMES]
for name, fn in (("single", lambda: eng.logp_grad("univariate", *hargs)),
                 ("overlapped4", lambda: eng.logp_grad_overlapped("univariate", *hargs, chunks=4))):
    ts = []
    for i in range(25):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3); del out
    print(name, " ".join(f"{t:.1f}" for t in ts))
