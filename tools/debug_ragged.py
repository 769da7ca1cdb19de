import sys, numpy as np, torch
sys.path.insert(0, '.')
from pymc_experimental_b200 import synth
from pymc_experimental_b200.engine import KalmanEngine, PARAM_NAMES as NAMES
from oracle import kalman_torch as oth
eng = KalmanEngine(0)
OUT = ("filtered_states","predicted_states","observed_states","filtered_covariances","predicted_covariances","observed_covariances","loglike_obs")
for kind in ("standard","univariate"):
  for B in (3,7):
    cfg = synth.random_dense(B, 17, 15, 1, 4, seed=40 + B, missing=0.2)
    rng = np.random.default_rng(B)
    cfg["P0"] = cfg["P0"] + 0.05 * rng.standard_normal(cfg["P0"].shape)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    out = eng.filter(kind, *args)
    ref = oth.kalman_filter(kind, *[torch.as_tensor(a) for a in args])
    print(kind, B, "missing y at:", np.argwhere(np.isnan(cfg["y"][...,0]))[:12].tolist())
    for n, r in zip(OUT, ref):
        g = np.asarray(out[n]).reshape(r.shape); r = r.numpy()
        bad = np.argwhere(np.isnan(g) != np.isnan(r))
        err = np.nanmax(np.abs(g - r)) / np.nanmax(np.abs(r))
        print("  ", n, "nan-mismatch:", bad[:4].tolist(), "nan in ref:", int(np.isnan(r).sum()), "nan in got:", int(np.isnan(g).sum()), "relerr %.2e" % err)
