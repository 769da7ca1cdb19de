import sys, numpy as np, torch
sys.path.insert(0, ".")
from pymc_experimental_b200 import synth
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine
eng = KalmanEngine(0)
cfg = synth.random_dense(296, 20, 32, 16, 16, seed=1, missing=0.0)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
args = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
for _ in range(2):
    eng.logp_grad("standard", *args)
torch.cuda.synchronize()
