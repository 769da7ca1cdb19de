"""Sharding of the independent (series x draw) batch over ranks and the single final gather.

One process per GPU; no traffic while filtering.  ``shard_bounds`` gives the contiguous block of the batch axis a
rank owns; ``gather_packed`` is the only collective of the path: one ``all_gather_into_tensor`` (NCCL over NVLink on
the GPU box, gloo in the CPU tests) of the packed ``[logp | g_a0 | ... | g_Q]`` buffer each rank produced
(``KalmanEngine.packed_layout``), followed by a local re-interleave into global batch order.
"""

from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block split: the first ``B % world`` ranks get one extra series."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_STATIC_NDIM = {"y": 2, "a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


def shard_inputs(arrays: dict, B: int, world: int, rank: int, shared=(), time_varying=(), batched=()):
    """Slice every batched array to this rank's block; the others are replicated.

    Whether an array carries the batch axis is decided by its RANK against the static rank of its name (``y`` [n, p],
    ``a0`` [m], ``P0`` [m, m] ... plus one for the names in ``time_varying``), as ``KalmanEngine`` does -- never by
    ``shape[0] == B``, which would silently slice a shared array whose leading dimension happens to equal B.  Arrays
    with other names are sliced only when listed in ``batched``; names in ``shared`` are never sliced."""
    lo, hi = shard_bounds(B, world, rank)
    out = {}
    for k, v in arrays.items():
        if k in shared:
            is_batched = False
        elif k in batched:
            is_batched = True
        elif k in _STATIC_NDIM:
            nd = _STATIC_NDIM[k] + int(k in time_varying)
            if v.ndim not in (nd, nd + 1):
                raise ValueError(f"{k}: expected {nd} (shared) or {nd + 1} (batched) dims, got {v.ndim}")
            is_batched = v.ndim == nd + 1
        else:
            raise ValueError(f"{k}: not a state-space input; list it in `batched` or `shared`")
        if is_batched and v.shape[0] != B:
            raise ValueError(f"{k}: leading batch dimension {v.shape[0]} != {B}")
        out[k] = v[lo:hi] if is_batched else v
    return out


def packed_sizes(m: int, p: int, r: int) -> dict:
    return {"logp": 1, "a0": m, "P0": m * m, "c": m, "d": p, "T": m * m, "Z": p * m, "R": m * r, "H": p * p,
            "Q": r * r}


def gather_packed(packed, B: int, m: int, p: int, r: int, group=None):
    """All-gather the per-rank packed buffers and return a dict of global arrays in batch order.

    ``packed`` is this rank's flat torch tensor of ``B_local * (1 + n_grad)`` elements laid out field-major
    (``[logp(B_local) | a0(B_local, m) | ...]``).  Ranks may own different ``B_local`` (block split), so buffers
    are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = packed_sizes(m, p, r)
    per_series = sum(sizes.values())
    blocks = [shard_bounds(B, world, rk) for rk in range(world)]
    bmax = max(hi - lo for lo, hi in blocks)
    send = packed.new_zeros(bmax * per_series)
    send[: packed.numel()] = packed
    recv = packed.new_empty(world * bmax * per_series)
    dist.all_gather_into_tensor(recv, send, group=group)
    shapes = {"logp": (), "a0": (m,), "P0": (m, m), "c": (m,), "d": (p,), "T": (m, m), "Z": (p, m), "R": (m, r),
              "H": (p, p), "Q": (r, r)}
    out = {k: packed.new_empty((B, *shapes[k])) for k in sizes}
    for rk, (lo, hi) in enumerate(blocks):
        bl = hi - lo
        chunk = recv[rk * bmax * per_series:]
        off = 0
        for k, sz in sizes.items():
            out[k][lo:hi] = chunk[off: off + bl * sz].reshape(bl, *shapes[k])
            off += bl * sz
    return out
