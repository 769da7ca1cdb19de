"""B = 1 latency of one logp + dlogp evaluation through the PyTensor Ops (what a PyMC model with the drop-in filter pays per
NUTS leapfrog step), C2 config (SARIMA m = 15, T = 500, UnivariateFilter), next to the fused engine call and the CPU port.

    python tools/bench_op_latency.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
try:
    import pytensor
except ImportError:
    sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
    import pytensor
import pytensor.tensor as pt  # noqa: E402

from pymc_experimental_b200 import filters, synth  # noqa: E402
from pymc_experimental_b200.engine import PARAM_NAMES, default_engine  # noqa: E402


def timeit(fn, reps=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    cfg = synth.sarima(B=1, n=500)
    ins = {k: cfg[k][0] for k in PARAM_NAMES}
    ins["y"] = cfg["y"]
    sv = {"y": pt.tensor("data", shape=(None, 1))}
    for k in PARAM_NAMES:
        sv[k] = pt.tensor(k, shape=ins[k].shape)
    outs = filters.UnivariateFilter().build_graph(sv["y"], *[sv[k] for k in PARAM_NAMES])
    wrt = [sv[k] for k in PARAM_NAMES]
    logp = outs[6].sum()
    f = pytensor.function([sv["y"]] + wrt, [logp, *pytensor.grad(logp, wrt)])
    vals = [ins["y"]] + [ins[k] for k in PARAM_NAMES]
    eng = default_engine(0)
    n0 = eng.launch_count
    f(*vals)
    launches = eng.launch_count - n0
    t_op = timeit(lambda: f(*vals))
    t_fused = timeit(lambda: eng.logp_grad("univariate", *vals))
    t_fwd = timeit(lambda: eng.forward_saved("univariate", *vals))
    from oracle import kalman_c

    kalman_c.lib()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    yb = ins["y"]
    prm = [ins[k][None] for k in PARAM_NAMES]
    t_cpu = timeit(lambda: kalman_c.logp_and_grad("univariate", yb, *prm), reps=50)
    print(json.dumps({
        "config": "C2 SARIMA m=15 T=500 UnivariateFilter, B=1, float64, host arrays in / out",
        "pytensor": pytensor.__version__,
        "op_logp_dlogp_ms": 1e3 * t_op, "kernel_launches_per_evaluation": launches,
        "engine_fused_logp_grad_ms": 1e3 * t_fused, "engine_forward_only_ms": 1e3 * t_fwd,
        "cpu_port_one_series_ms": 1e3 * t_cpu, "cpu_port": "oracle/kalman_c.c, one series on one core",
        "state_steps_per_s_op": 500 / t_op, "state_steps_per_s_cpu_port": 500 / t_cpu,
    }))


if __name__ == "__main__":
    main()
