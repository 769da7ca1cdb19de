import sys, numpy as np, torch, json
sys.path.insert(0, ".")
from pymc_experimental_b200 import synth
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine
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


for (m, r, B, n) in ((15, 1, 4096, 200), (17, 1, 4096, 200), (27, 1, 4096, 200), (40, 2, 2048, 200)):
    cfg = synth.random_dense(B, n, m, 1, r, seed=1, missing=0.0)
    args = [dev(cfg["y"])] + [dev(cfg[k]) for k in PARAM_NAMES]
    for kind in ("standard", "univariate"):
        ms = timeit(lambda: eng.logp_grad(kind, *args))
        print({"m": m, "p": 1, "B": B, "T": n, "filter": kind, "path": eng.last_path, "ms": round(ms, 2), "state_steps_per_s": round(B * n / ms * 1e3)}, flush=True)
