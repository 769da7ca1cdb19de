import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import kalman_torch as oth
from pymc_experimental_b200 import synth
from pymc_experimental_b200.engine import KalmanEngine
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
eng = KalmanEngine(0)
for eps in (1e-2, 1e-3, 1e-5):
    m, p, r = 16, 8, 8
    B, n = 6, 12
    cfg = synth.random_dense(B, n, m, p, r, seed=77 + m, missing=0.0)
    rng = np.random.default_rng(5)
    L = np.linalg.cholesky(cfg["P0"]) @ np.linalg.qr(rng.standard_normal((B, m, m)))[0]
    L[:, 1, :] = L[:, 0, :] + eps * np.abs(L[:, 0, :]).max() * rng.standard_normal((B, m))
    L[:, 11, :] = L[:, 9, :] + eps * np.abs(L[:, 9, :]).max() * rng.standard_normal((B, m))
    cfg["P0"] = L
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("cholesky", *args)
    ref_logp, ref_g = oth.logp_and_grad("cholesky", *args)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    print(eps, eng.last_path, "logp", rel(logp, ref_logp), {k: f"{rel(grads[k], ref_g[k]):.1e}" for k in NAMES}, "max|gP0| gpu/ref", np.abs(grads["P0"]).max(), np.abs(ref_g["P0"]).max(), "cond", np.linalg.cond(L[0]))
