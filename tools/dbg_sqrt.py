"""Debug aid for the SquareRootFilter kernels (warp-per-series forward sweep, four-warp adjoint).

    python tools/dbg_sqrt.py time       C4-sized timing (DBG_B, DBG_T, DBG_MODE=fwd|grad)
    python tools/dbg_sqrt.py profile    clock64 phase timers of the adjoint (a BKF_SQC_PROFILE build of the library:
                                        python tools/build_variant.py prof "-DBKF_SQC_PROFILE" bkf_sqrt_bwd4_inst_a.cu ;
                                        DBG_LIB=libbkf_prof.so python tools/dbg_sqrt.py profile)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")


def timing():
    sys.path.insert(0, ROOT)
    import torch

    from pymc_experimental_b200 import engine as _eng
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

    if os.environ.get("DBG_LIB"):
        _eng.LIB_PATH = os.path.join(os.path.dirname(_eng.LIB_PATH), os.environ["DBG_LIB"])
    eng = KalmanEngine(0)
    B, n = int(os.environ.get("DBG_B", 512)), int(os.environ.get("DBG_T", 200))
    cfg = synth.varmax(B, n)
    dev = torch.device("cuda", 0)
    dargs = [torch.as_tensor(cfg["y"], device=dev)] + [torch.as_tensor(cfg[k], device=dev) for k in NAMES]
    mode = os.environ.get("DBG_MODE", "fwd")
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "fwd":
            res = eng.filter("cholesky", *dargs, outputs=("loglike_obs",))
            tot = float(res["loglike_obs"].sum())
        else:
            logp, grads, status = eng.logp_grad("cholesky", *dargs)
            tot = float(logp.sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{mode} B={B} T={n} path={eng.last_path} {dt*1e3:.2f} ms sum={tot:.12e}", flush=True)


def profile():
    """BKF_SQC_PROFILE build: phase timers of the four-warp adjoint (bkf_sqrt_bwd4.cuh), cycles per step, per warp."""
    sys.path.insert(0, ROOT)
    from pymc_experimental_b200 import engine as _eng
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

    if os.environ.get("DBG_LIB"):
        _eng.LIB_PATH = os.path.join(os.path.dirname(_eng.LIB_PATH), os.environ["DBG_LIB"])
    eng = KalmanEngine(0)
    B, n = int(os.environ.get("DBG_B", 512)), 100
    cfg = synth.varmax(B, n)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    for _ in range(2):
        logp, grads, status = eng.logp_grad("cholesky", *args)
    t = grads["T"][0].reshape(-1)[64:128] / n
    names = ["S0p,v,turn", "pass1p,inv", "pass2p,TPt", "vecF|U1t", "gT,Pcfb,afb", "vecU|S0a", "S0b", "pass1u", "pass2u", "A21,HB", "rows",
             "gZ,abar"]
    print("B", B, eng.last_path)
    for w in range(4):
        print("warp", w, {k: int(v) for k, v in zip(names, t[16 * w:16 * w + 12])}, "sum", int(t[16 * w:16 * w + 16].sum()))


if __name__ == "__main__":
    {"time": timing, "profile": profile}[sys.argv[1]]()
