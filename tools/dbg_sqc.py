"""Debug aid: four-warp square-root kernels (BKF_SQRT_CTA) against the warp-per-series ones, outputs and timing.

    python tools/dbg_sqc.py fwd        forward outputs of every instantiated shape, with and without missing values
    python tools/dbg_sqc.py grad       logp + gradients
    python tools/dbg_sqc.py time       C4-sized timing of both families
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
OUTS = ("filtered_states", "predicted_states", "observed_states", "filtered_covariances", "predicted_covariances",
        "observed_covariances", "loglike_obs")
SHAPES = [(16, 8, 8), (24, 8, 8), (16, 16, 16), (32, 16, 16)]


def make(shape, missing, B=9, n=14):
    from pymc_experimental_b200 import synth

    m, p, r = shape
    cfg = synth.random_dense(B, n, m, p, r, seed=300 + m, missing=0.0)
    if missing:
        rng = np.random.default_rng(m)
        cfg["y"][rng.uniform(size=(B, n)) < missing] = np.nan  # whole observation vectors
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    return [cfg["y"]] + [cfg[k] for k in NAMES]


def child(what, shape, missing, out):
    sys.path.insert(0, ROOT)
    from pymc_experimental_b200.engine import KalmanEngine

    eng = KalmanEngine(0)
    args = make(shape, missing)
    if what == "fwd":
        res = eng.filter("cholesky", *args)
        np.savez(out, path=eng.last_path, **res)
    else:
        logp, grads, status = eng.logp_grad("cholesky", *args)
        np.savez(out, logp=logp, status=status, path=eng.last_path, **{"g_" + k: grads[k] for k in NAMES})


def timing():
    sys.path.insert(0, ROOT)
    import torch

    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

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
        print(f"BKF_SQRT_CTA={os.environ.get('BKF_SQRT_CTA')} {mode} B={B} T={n} path={eng.last_path} {dt*1e3:.2f} ms sum={tot:.12e}", flush=True)


def profile():
    """BKF_SQC_PROFILE build: phase timers of series 0 (cycles per step, per warp)."""
    sys.path.insert(0, ROOT)
    from pymc_experimental_b200 import engine as _eng
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

    if os.environ.get("DBG_LIB"):  # profiling variants of the library built beside the product one
        _eng.LIB_PATH = os.path.join(os.path.dirname(_eng.LIB_PATH), os.environ["DBG_LIB"])
    eng = KalmanEngine(0)
    B, n = int(os.environ.get("DBG_B", 512)), 100
    cfg = synth.varmax(B, n)
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    print("B", B)
    res = eng.filter("cholesky", *args, outputs=("loglike_obs",))
    res = eng.filter("cholesky", *args, outputs=("loglike_obs",))
    t = res["loglike_obs"][0][:48] / n
    names = ["S1", "B1wait", "S2 qr", "B2wait", "S3", "B3wait", "S4", "B0wait"]
    for w in range(4):
        print("warp", w, {k: int(v) for k, v in zip(names, t[8 * w:8 * w + 8])}, "sum", int(t[8 * w:8 * w + 8].sum()))
    npan = {1: 1, 2: 1, 3: 2, 4: 2, 5: 2, 6: 2}
    print("chain warp: HAND wait/step", int(t[32]), "syncs/step", t[33], "cycles per reflector by NQ:",
          {q: int(t[34 + q] / (8 * npan[q])) for q in range(1, 7)}, "panels total/step", int(t[35:41].sum()))


def profile_bwd():
    """BKF_SQC_PROFILE build: phase timers of the adjoint sweep of series 0 (cycles per step, per warp)."""
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
    names = ["load", "PS+ldRf", "solve|TPt", "U1t", "U1", "gT,Pcfb", "flag", "vec solves", "Rbar", "A21..rows", "gZ,abar",
             "-", "-", "-", "psi4", "psi6"]
    print("B", B, eng.last_path)
    for w in range(4):
        print("warp", w, {k: int(v) for k, v in zip(names, t[16 * w:16 * w + 16]) if k != "-"}, "sum", int(t[16 * w:16 * w + 16].sum()))


if __name__ == "__main__":
    if sys.argv[1] == "profile_bwd":
        profile_bwd()
        sys.exit(0)
    if sys.argv[1] == "profile":
        profile()
        sys.exit(0)
    if sys.argv[1] == "child":
        child(sys.argv[2], tuple(json.loads(sys.argv[3])), float(sys.argv[4]), sys.argv[5])
        sys.exit(0)
    what = sys.argv[1]
    if what == "time":
        timing()
        sys.exit(0)
    bit = "1" if what == "fwd" else "2"
    bad = 0
    for shape in SHAPES:
        for missing in (0.0, 0.2):
            res = {}
            for mode in ("0", bit):
                out = f"/tmp/sqc_{mode}.npz"
                env = dict(os.environ, BKF_SQRT_CTA=mode)
                r = subprocess.run([sys.executable, __file__, "child", what, json.dumps(shape), str(missing), out], env=env,
                                   timeout=120)
                if r.returncode != 0:
                    print(shape, missing, "mode", mode, "FAILED rc", r.returncode, flush=True)
                    res = None
                    break
                res[mode] = np.load(out)
            if res is None:
                bad += 1
                continue
            a, b = res["0"], res[bit]
            keys = OUTS if what == "fwd" else ("logp",) + tuple("g_" + k for k in NAMES)
            worst = {}
            for k in keys:
                sc = max(np.nanmax(np.abs(a[k])), 1e-300)
                diff = np.abs(a[k] - b[k])
                diff[np.isnan(a[k]) & np.isnan(b[k])] = 0.0
                worst[k] = float(np.nanmax(diff) / sc) if not np.isnan(diff).any() else float("nan")
            w = max(worst.values())
            ok = w < 1e-9 and (a["status"] == b["status"]).all()
            bad += 0 if ok else 1
            print(shape, missing, str(a["path"]), "->", str(b["path"]), "worst rel", f"{w:.2e}", "OK" if ok else "MISMATCH",
                  {k: f"{v:.1e}" for k, v in worst.items() if not v < 1e-9}, flush=True)
    print("bad", bad)
    sys.exit(1 if bad else 0)
