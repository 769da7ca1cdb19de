"""bkf_comm_* across two processes / two GPUs: each rank filters its block of the batch and the packed [logp | grads]
buffers are gathered by the library's own NCCL call (no torch.distributed on the data path; a file carries the
communicator id).  Skipped unless two GPUs are visible (gpurun --gpus 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")

CHILD = r"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, {root!r})
from pymc_experimental_b200 import synth, sharding
from pymc_experimental_b200.engine import KalmanEngine, PARAM_NAMES
rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
eng = KalmanEngine(rank)
if rank == 0:
    open(idfile + ".tmp", "wb").write(KalmanEngine.comm_unique_id()); os.replace(idfile + ".tmp", idfile)
while not os.path.exists(idfile): time.sleep(0.05)
eng.comm_init(world, rank, open(idfile, "rb").read())
B, n = 24, 40
cfg = synth.sarima(B=B, n=n)
lo, hi = sharding.shard_bounds(B, world, rank)
dev = torch.device("cuda", rank)
torch.cuda.set_device(dev)
dargs = [torch.as_tensor(cfg["y"], device=dev)] + [torch.as_tensor(cfg[k][lo:hi], device=dev) for k in PARAM_NAMES]
layout, total = eng.packed_layout(hi - lo, 15, 1, 1)
packed = torch.empty(total, dtype=torch.float64, device=dev)
eng.logp_grad("univariate", *dargs, packed_out=packed)
g = eng.comm_allgather(packed)
torch.cuda.synchronize()
np.save(out, g.cpu().numpy())
eng.comm_destroy()
"""


def test_two_rank_gather_through_the_c_abi(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pymc_experimental_b200 import synth
    from pymc_experimental_b200.engine import KalmanEngine

    script = tmp_path / "child.py"
    script.write_text(CHILD.format(root=ROOT))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, str(tmp_path / f"out{r}.npy")])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    g0, g1 = np.load(tmp_path / "out0.npy"), np.load(tmp_path / "out1.npy")
    np.testing.assert_array_equal(g0, g1)
    # the gathered buffer is the two ranks' packed buffers: compare with the whole batch on one GPU
    eng = KalmanEngine(0)
    cfg = synth.sarima(B=24, n=40)
    logp, grads, _ = eng.logp_grad("univariate", cfg["y"], *[cfg[k] for k in NAMES])
    layout, total = eng.packed_layout(12, 15, 1, 1)
    off, size, _ = layout["logp"]
    both = np.concatenate([g0[r, off:off + size] for r in range(2)])
    np.testing.assert_array_equal(both, logp)
    off, size, _ = layout["T"]
    gT = np.concatenate([g0[r, off:off + size].reshape(12, 15, 15) for r in range(2)])
    np.testing.assert_array_equal(gT, grads["T"])
