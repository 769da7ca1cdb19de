"""world_size-2 gloo test (CPU) of the only multi-rank logic of the path: block sharding of the batch axis and the
single gather of the packed [logp | grads] buffer.  The per-rank compute is stood in for by the oracle (the CUDA
engine cannot run here); the packing layout is the engine's own (KalmanEngine.packed_layout)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pymc_experimental_b200 import sharding, synth
from pymc_experimental_b200.engine import PARAM_NAMES, KalmanEngine

B, N, M, P, R = 7, 9, 3, 1, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    from oracle import kalman_torch as oth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.random_dense(B, N, M, P, R, seed=5)
    local = sharding.shard_inputs(cfg, B, world, rank)
    lo, hi = sharding.shard_bounds(B, world, rank)
    logp, grads = oth.logp_and_grad("standard", local["y"], *[local[k] for k in PARAM_NAMES])
    layout, total = KalmanEngine.packed_layout(hi - lo, M, P, R)
    packed = torch.empty(total, dtype=torch.float64)
    packed[layout["logp"][0]: layout["logp"][0] + layout["logp"][1]] = torch.as_tensor(logp)
    for k in PARAM_NAMES:
        off, sz, _ = layout[k]
        packed[off: off + sz] = torch.as_tensor(grads[k]).reshape(-1)
    out = sharding.gather_packed(packed, B, M, P, R)
    if rank == 0:
        ret.put({k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for b in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [sharding.shard_bounds(b, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == b
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gather_equals_single_rank():
    from oracle import kalman_torch as oth

    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = ret.get(timeout=240)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    cfg = synth.random_dense(B, N, M, P, R, seed=5)
    logp, grads = oth.logp_and_grad("standard", cfg["y"], *[cfg[k] for k in PARAM_NAMES])
    np.testing.assert_allclose(got["logp"], logp, rtol=1e-13)
    for k in PARAM_NAMES:
        np.testing.assert_allclose(got[k], grads[k], rtol=1e-12, atol=1e-14)
