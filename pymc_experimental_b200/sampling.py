"""Conditional-posterior draws from the smoother's output.

Host-side mirror of ``SequenceMvNormal`` (pymc_extras/statespace/filters/distributions.py:411-443): the reference scans
``MvNormalSVD(mu=mu[t], cov=cov[t])`` over the time axis to sample hidden states given the smoothed moments
(``core/statespace.py:1119-1150``).  Here every (series, step) is drawn in one kernel launch: the standard normal
variates come from the caller's generator (NumPy ``Generator`` on the host, ``torch.Generator`` on the device) or --
``seed=`` -- from the kernels' own counter-based generator (Philox4x32-10 + Box-Muller; no variates cross PCIe), the
square root of each covariance and the affine map run on the GPU (``bkf_mvn_sample``).

The draws have the reference's DISTRIBUTION, not its numbers: NumPy's SVD method uses the square root
``sqrt(s) V^T``, the kernel a Cholesky factor with semidefinite pivots dropped, and the two differ by an orthogonal
transformation of the variates (a different random stream gives different numbers in the reference as well).
"""

from __future__ import annotations

import numpy as np

from .engine import KalmanEngine, _is_torch


def sample_sequence_mvnormal(mus, covs, rng=None, engine: KalmanEngine | None = None, return_status=False, seed=None,
                             offset=0):
    """One draw per (batch..., time) from ``N(mus[..., t, :], covs[..., t, :, :])``.

    ``mus``: (..., T, m), ``covs``: (..., T, m, m), NumPy arrays or CUDA tensors (the result lives where the inputs do).
    ``rng``: ``numpy.random.Generator`` (host inputs) or ``torch.Generator`` (device inputs); a fresh default one if None.
    ``seed`` (int): draw the variates in the kernel instead (``offset``: counter to start from).
    """
    eng = engine if engine is not None else KalmanEngine(0)
    if seed is not None:
        lead = tuple(mus.shape[:-2])
        x, status = eng.sample_mvn(mus.reshape((-1,) + tuple(mus.shape[-2:])),
                                   covs.reshape((-1,) + tuple(covs.shape[-3:])), seed=seed, offset=offset)
        if not _is_torch(mus):
            x, status = np.array(x), np.array(status)
        x = x.reshape(tuple(mus.shape))
        status = status.reshape(lead) if lead else status[0]
    elif _is_torch(mus):
        import torch

        z = torch.randn(mus.shape, dtype=mus.dtype, device=mus.device, generator=rng)
        lead = tuple(mus.shape[:-2])
        x, status = eng.sample_mvn(mus.reshape((-1,) + tuple(mus.shape[-2:])),
                                   covs.reshape((-1,) + tuple(covs.shape[-3:])), z.reshape((-1,) + tuple(mus.shape[-2:])))
        x = x.reshape(tuple(mus.shape))
        status = status.reshape(lead) if lead else status[0]
    else:
        mus, covs = np.asarray(mus), np.asarray(covs)
        gen = rng if rng is not None else np.random.default_rng()
        z = gen.standard_normal(mus.shape).astype(mus.dtype, copy=False)
        lead = mus.shape[:-2]
        x, status = eng.sample_mvn(mus.reshape((-1,) + mus.shape[-2:]), covs.reshape((-1,) + covs.shape[-3:]),
                                   z.reshape((-1,) + mus.shape[-2:]))
        x = np.array(x).reshape(mus.shape)
        status = np.array(status).reshape(lead) if lead else int(status[0])
    return (x, status) if return_status else x


def simulate_statespace(a0, P0, c, d, T, Z, R, H, Q, steps, draws=None, rng=None, engine: KalmanEngine | None = None,
                        append_x0=True, sequence_names=(), seed=None, offset=0):
    """Unconditional draws of the hidden states and observations: the batched ``LinearGaussianStateSpace.rv_op``
    (filters/distributions.py:181-293).  Any of the nine matrices may carry a leading batch axis (one model per draw);
    ``draws`` gives the batch size when none does.  Returns an array (B, steps + 1, m + p) -- states and observations
    concatenated along the last axis as the reference does (``steps`` rows without the initial draw if not ``append_x0``).
    ``sequence_names``: matrices that carry a time axis of length ``steps`` right after the optional batch axis.
    """
    eng = engine if engine is not None else KalmanEngine(0)
    mats = dict(zip(("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q"), (a0, P0, c, d, T, Z, R, H, Q)))
    static = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}
    B = None
    for k, v in mats.items():
        if v.ndim == static[k] + 1 + int(k in sequence_names):
            B = int(v.shape[0])
    if B is None:
        B = 1 if draws is None else int(draws)
    m, r, p = int(mats["R"].shape[-2]), int(mats["R"].shape[-1]), int(mats["Z"].shape[-2])
    if seed is not None:  # variates drawn in the kernel
        on_dev = _is_torch(mats["T"])
        xs, ys, _ = eng.simulate(*(mats.values() if on_dev else [np.asarray(v) for v in mats.values()]), batch=B,
                                 steps=steps, seed=seed, offset=offset, time_varying=tuple(sequence_names))
        if on_dev:
            import torch

            out = torch.cat([xs, ys], dim=-1)
        else:
            out = np.concatenate([np.array(xs), np.array(ys)], axis=-1)
    elif _is_torch(mats["T"]):
        import torch

        kw = dict(dtype=mats["T"].dtype, device=mats["T"].device, generator=rng)
        zs = [torch.randn(sh, **kw) for sh in ((B, m), (B, p), (B, steps, r), (B, steps, p))]
        xs, ys, _ = eng.simulate(*mats.values(), *zs, time_varying=tuple(sequence_names))
        out = torch.cat([xs, ys], dim=-1)
    else:
        gen = rng if rng is not None else np.random.default_rng()
        zs = [gen.standard_normal(sh) for sh in ((B, m), (B, p), (B, steps, r), (B, steps, p))]
        xs, ys, _ = eng.simulate(*[np.asarray(v) for v in mats.values()], *zs, time_varying=tuple(sequence_names))
        out = np.concatenate([np.array(xs), np.array(ys)], axis=-1)
    return out if append_x0 else out[:, 1:]


def sample_conditional_posterior(kind, data, a0, P0, c, d, T, Z, R, H, Q, rng=None, engine: KalmanEngine | None = None,
                                 time_varying=(), seed=None, offset=0):
    """Hidden-state draws given the data for a batch of parameter draws: filter, smoother and one
    ``N(smoothed_state, smoothed_cov)`` draw per (draw, step) in one library call (``bkf_filter_smooth_sample``) -- what
    ``PyMCStateSpace.sample_conditional_posterior`` (core/statespace.py:1061-1171) evaluates draw by draw.  Returns
    ``(x (B, T, m), loglike_obs (B, T))``; only those cross the host-device boundary."""
    eng = engine if engine is not None else KalmanEngine(0)
    m = int(R.shape[-2])
    n = int(data.shape[-2])
    static = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}
    mats = dict(zip(static, (a0, P0, c, d, T, Z, R, H, Q)))
    B = int(data.shape[0]) if data.ndim == 3 else None
    for k, v in mats.items():
        if v.ndim == static[k] + 1 + int(k in time_varying):
            B = int(v.shape[0])
    shape = (n, m) if B is None else (B, n, m)
    if seed is not None:  # variates drawn in the kernel: only x and loglike_obs cross the host-device boundary
        x, ll, _ = eng.sample_conditional_posterior(kind, data, *mats.values(), time_varying=tuple(time_varying),
                                                    seed=seed, offset=offset)
        return x, ll
    if _is_torch(data):
        import torch

        z = torch.randn(shape, dtype=data.dtype, device=data.device, generator=rng)
    else:
        z = (rng if rng is not None else np.random.default_rng()).standard_normal(shape)
    x, ll, _ = eng.sample_conditional_posterior(kind, data, *mats.values(), z, time_varying=tuple(time_varying))
    return x, ll
