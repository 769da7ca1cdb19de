"""Debug aid: warp-per-series square-root kernels against the CTA-per-series ones (run twice with BKF_SQRT_WARP)."""
import os, subprocess, sys, json
import numpy as np

NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")

def run(shape, missing, out):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine
    eng = KalmanEngine(0)
    m, p, r = shape
    B, n = 9, 14
    cfg = synth.random_dense(B, n, m, p, r, seed=300 + m, missing=0.0)
    if missing:
        rng = np.random.default_rng(m)
        cfg["y"][rng.uniform(size=(B, n)) < missing] = np.nan
    cfg["P0"] = np.linalg.cholesky(cfg["P0"])
    args = [cfg["y"]] + [cfg[k] for k in NAMES]
    logp, grads, status = eng.logp_grad("cholesky", *args)
    np.savez(out, logp=logp, status=status, path=eng.last_path, **{"g_" + k: grads[k] for k in NAMES})

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run(tuple(json.loads(sys.argv[2])), float(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    modes = sys.argv[1:] or ["1"]
    for shape in [(8, 8, 8), (16, 8, 8), (24, 8, 8), (32, 8, 8), (16, 16, 8), (16, 16, 16), (32, 16, 16)]:
        for missing in (0.0, 0.2):
            res = {}
            for mode in ["0"] + modes:
                out = f"/tmp/sqw_{mode}.npz"
                env = dict(os.environ, BKF_SQRT_WARP=mode)
                subprocess.run([sys.executable, __file__, "child", json.dumps(shape), str(missing), out], env=env, check=True)
                res[mode] = np.load(out)
            for mode in modes:
                a, b = res["0"], res[mode]
                worst = abs(a["logp"] - b["logp"]).max() / abs(a["logp"]).max()
                gw = {}
                for k in NAMES:
                    sc = max(abs(a["g_" + k]).max(), 1e-300)
                    gw[k] = float(abs(a["g_" + k] - b["g_" + k]).max() / sc)
                print(shape, missing, "mode", mode, str(b["path"]), "logp rel", f"{worst:.2e}", "grad rel max", f"{max(gw.values()):.2e}",
                      {k: f"{v:.1e}" for k, v in gw.items() if v > 1e-9}, flush=True)
