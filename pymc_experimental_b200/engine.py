"""ctypes binding of libbkf.so (include/bkf.h) -- the only way into the CUDA kernels from Python.

There is no CPU fallback: if the shared library is missing or no CUDA device is present, constructing a
:class:`KalmanEngine` raises.  Arrays may be NumPy arrays (host pointers; the library stages them through its
device workspace) or ``torch`` CUDA tensors (device pointers, used in place; outputs are returned as torch
tensors on the same device).
"""

from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("BKF_LIB", "libbkf.so"))  # BKF_LIB: a tools/build_variant.py build

STANDARD, UNIVARIATE, CHOLESKY = 0, 1, 2
KINDS = {"standard": STANDARD, "univariate": UNIVARIATE, "cholesky": CHOLESKY}
F64, F32 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
PARAM_NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")  # utils/constants.py:24 (x0 == a0)
STATUS_NOT_PD, STATUS_NONFINITE, STATUS_SMOOTHER_ILL, STATUS_NOT_STATIONARY = 1, 2, 4, 8

# pymc_extras/statespace/utils/constants.py:19-20
MISSING_FILL = -9999.0
JITTER_F64, JITTER_F32 = 1e-8, 1e-6

# every symbol include/bkf.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = (
    "bkf_version", "bkf_create", "bkf_destroy", "bkf_last_error", "bkf_set_stream", "bkf_launch_count",
    "bkf_last_path", "bkf_enable_timing", "bkf_reset_timing", "bkf_kernel_time_ms", "bkf_measure_fma_peak",
    "bkf_filter_forward", "bkf_filter_logp_grad", "bkf_smooth", "bkf_filter_smooth", "bkf_host_alloc", "bkf_host_free",
    "bkf_logp_grad_workspace_bytes", "bkf_device_memory", "bkf_filter_logp_grad_compact", "bkf_stationary_init",
    "bkf_stationary_init_grad", "bkf_mvn_sample", "bkf_simulate",
    "bkf_filter_smooth_sample", "bkf_filter_forward_saved", "bkf_filter_backward_saved", "bkf_filter_logp_grad_theta",
    "bkf_comm_unique_id", "bkf_comm_init", "bkf_comm_allgather", "bkf_comm_destroy",
)
COMM_ID_BYTES = 128
FN_ID, FN_COS, FN_SIN, FN_EXP, FN_ONE_MINUS, FN_SQUARE = range(6)  # BKF_FN_* (theta -> entries programs)
ERR_STALE = -5


class BkfProblem(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("dtype", C.c_int), ("mem", C.c_int), ("B", C.c_int), ("n", C.c_int),
        ("m", C.c_int), ("p", C.c_int), ("r", C.c_int), ("y_shared", C.c_int),
        ("shared_mask", C.c_uint), ("tv_mask", C.c_uint),
        ("cov_jitter", C.c_double), ("missing_fill", C.c_double),
        ("rng_seed", C.c_ulonglong), ("rng_offset", C.c_ulonglong),
    ]


class BkfError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libbkf.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BkfError(
            f"{LIB_PATH} not found: build it with `python -m pymc_experimental_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    lib.bkf_version.restype = C.c_int
    lib.bkf_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.bkf_destroy.argtypes = [vp]
    lib.bkf_last_error.argtypes = [vp]
    lib.bkf_last_error.restype = C.c_char_p
    lib.bkf_last_path.argtypes = [vp]
    lib.bkf_last_path.restype = C.c_char_p
    lib.bkf_set_stream.argtypes = [vp, vp]
    lib.bkf_launch_count.argtypes = [vp]
    lib.bkf_launch_count.restype = C.c_longlong
    lib.bkf_enable_timing.argtypes = [vp, C.c_int]
    lib.bkf_reset_timing.argtypes = [vp]
    lib.bkf_kernel_time_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.bkf_measure_fma_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.bkf_logp_grad_workspace_bytes.argtypes = [C.POINTER(BkfProblem), C.POINTER(C.c_size_t)]
    lib.bkf_device_memory.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.bkf_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.bkf_host_free.argtypes = [vp]
    pp = C.POINTER(BkfProblem)
    lib.bkf_filter_forward.argtypes = [vp, pp] + [vp] * 10 + [vp] * 7 + [vp]
    lib.bkf_filter_logp_grad.argtypes = [vp, pp] + [vp] * 10 + [vp] + [vp] + [vp] * 9 + [vp]
    lib.bkf_filter_forward_saved.argtypes = [vp, pp] + [vp] * 10 + [vp, vp, vp, C.POINTER(C.c_ulonglong)]
    lib.bkf_filter_backward_saved.argtypes = [vp, C.c_ulonglong, vp] + [vp] * 9
    lib.bkf_filter_logp_grad_compact.argtypes = [vp, pp, vp, C.POINTER(vp), C.c_int, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int), vp, vp, vp, vp, vp]
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.bkf_filter_logp_grad_theta.argtypes = [vp, pp, vp, C.POINTER(vp), C.c_int, ip, ip, C.c_int, ip, dp, ip, ip, ip,
                                               C.c_int, vp, vp, vp, vp, vp]
    lib.bkf_stationary_init.argtypes = [vp, pp] + [vp] * 4 + [vp] * 2 + [vp]
    lib.bkf_stationary_init_grad.argtypes = [vp, pp] + [vp] * 7 + [vp] * 4
    lib.bkf_mvn_sample.argtypes = [vp, pp] + [vp] * 3 + [vp] + [vp]
    lib.bkf_simulate.argtypes = [vp, pp] + [vp] * 9 + [vp] * 4 + [vp] * 2 + [vp]
    lib.bkf_filter_smooth_sample.argtypes = [vp, pp] + [vp] * 10 + [vp] + [vp] * 2 + [vp]
    lib.bkf_smooth.argtypes = [vp, pp] + [vp] * 5 + [vp] * 2 + [vp]
    lib.bkf_filter_smooth.argtypes = [vp, pp] + [vp] * 10 + [vp] * 3 + [vp]
    lib.bkf_comm_unique_id.argtypes = [vp]
    lib.bkf_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.bkf_comm_allgather.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    lib.bkf_comm_destroy.argtypes = [vp]
    _lib = lib
    return lib


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _PinnedPool:
    """Page-locked host blocks for call outputs (bkf_host_alloc), recycled when the NumPy arrays built on them die.

    D2H copies into pageable ``np.empty`` arrays are staged by the driver at a fraction of the link rate; outputs
    handed out by this pool are ordinary ndarrays (safe to keep: a block returns to the pool only after every view
    of it has been garbage-collected)."""

    GRAIN = 1 << 16
    MAX_FREE_BYTES = 4 << 30

    def __init__(self, lib):
        self.lib = lib
        self.free = {}
        self.free_bytes = 0

    def empty(self, shape, dtype):
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        if nbytes < self.GRAIN:  # small outputs: pageable memory is fine
            return np.empty(shape, dtype=dtype)
        size = -(-nbytes // self.GRAIN) * self.GRAIN
        blocks = self.free.get(size)
        if blocks:
            ptr = blocks.pop()
            self.free_bytes -= size
        else:
            p = C.c_void_p()
            if self.lib.bkf_host_alloc(size, C.byref(p)) != 0 or not p.value:
                return np.empty(shape, dtype=dtype)
            ptr = p.value
        cbuf = (C.c_char * size).from_address(ptr)
        weakref.finalize(cbuf, self._release, ptr, size)
        return np.frombuffer(cbuf, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize).reshape(shape)

    def _release(self, ptr, size):
        if self.free_bytes + size > self.MAX_FREE_BYTES:
            self.lib.bkf_host_free(C.c_void_p(ptr))
            return
        self.free.setdefault(size, []).append(ptr)
        self.free_bytes += size


_pinned_pool = None


def pinned_pool():
    global _pinned_pool
    if _pinned_pool is None:
        _pinned_pool = _PinnedPool(load_library())
    return _pinned_pool


class _Buffers:
    """Normalises one call's arrays to a single memory space / dtype and hands out raw pointers."""

    def __init__(self, arrays, dtype=None):
        self.device = any(_is_torch(a) and a.is_cuda for a in arrays if a is not None)
        if self.device:
            import torch

            self.torch = torch
            first = next(a for a in arrays if a is not None and _is_torch(a) and a.is_cuda)
            self.dev = first.device
            self.tdtype = first.dtype if dtype is None else dtype
            if self.tdtype not in (torch.float64, torch.float32):
                raise BkfError("dtype must be float64 or float32")
            self.code = F64 if self.tdtype == torch.float64 else F32
        else:
            nd = np.dtype(np.float64 if dtype is None else dtype)
            if nd not in (np.dtype(np.float64), np.dtype(np.float32)):
                raise BkfError("dtype must be float64 or float32")
            self.ndtype = nd
            self.code = F64 if nd == np.float64 else F32
        self.keep = []

    def inp(self, x):
        if self.device:
            t = self.torch.as_tensor(x, dtype=self.tdtype, device=self.dev).contiguous()
            self.keep.append(t)
            return t
        a = np.ascontiguousarray(x, dtype=self.ndtype)
        self.keep.append(a)
        return a

    def out(self, shape):
        if self.device:
            t = self.torch.empty(shape, dtype=self.tdtype, device=self.dev)
        else:
            t = pinned_pool().empty(shape, self.ndtype)
        self.keep.append(t)
        return t

    def out_int(self, shape):
        if self.device:
            t = self.torch.zeros(shape, dtype=self.torch.int32, device=self.dev)
        else:
            t = np.zeros(shape, dtype=np.int32)
        self.keep.append(t)
        return t

    @staticmethod
    def ptr(x):
        if x is None:
            return None
        if _is_torch(x):
            return C.c_void_p(x.data_ptr())
        return C.c_void_p(x.ctypes.data)


_STATIC_NDIM = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}


class KalmanEngine:
    """One handle = one device + one workspace + one stream (not thread-safe; create one per process)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.bkf_create(int(device), C.byref(h))
        if rc != 0:
            raise BkfError(f"bkf_create(device={device}) failed with code {rc}: no usable CUDA device "
                           "(this engine has no CPU fallback)")
        self.h = h
        self.device = int(device)
        self._stream = 0
        self._saved = None  # (ticket, buffers kept alive, shapes) of the last forward_saved
        # logp_grad splits a batch whose device workspace (trajectory + staging) would exceed this many bytes;
        # None = 70 % of the device's memory (B200: ~125 of 180 GB)
        self.max_workspace_bytes = None

    def device_memory(self):
        """(free, total) bytes of this engine's device."""
        f, t = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bkf_device_memory(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    # ---- the one collective of the path through the C ABI (bkf_comm_*): for callers without torch.distributed ----
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes that rank 0 hands to every rank (``ncclGetUniqueId``)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load_library().bkf_comm_unique_id(buf)
        if rc != 0:
            raise BkfError(f"bkf_comm_unique_id failed ({rc}): is libnccl.so.2 installed?")
        return buf.raw

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        if len(unique_id) != COMM_ID_BYTES:
            raise BkfError(f"unique_id must be {COMM_ID_BYTES} bytes")
        self._check(self.lib.bkf_comm_init(self.h, int(world), int(rank), C.c_char_p(unique_id)))
        self._comm_world = int(world)

    def comm_allgather(self, send, out=None):
        """All-gather of a contiguous device tensor over the communicator of ``comm_init`` (enqueued on the tensor's
        current stream behind the kernels that wrote it); returns a tensor with a leading axis of length world."""
        import torch

        if not send.is_cuda or send.device.index != self.device or not send.is_contiguous():
            raise BkfError("comm_allgather takes a contiguous tensor on the engine's device")
        code = {torch.float64: F64, torch.float32: F32}.get(send.dtype)
        if code is None:
            raise BkfError("comm_allgather: float64 or float32")
        world = getattr(self, "_comm_world", 0)
        if out is None:
            out = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        stream = torch.cuda.current_stream(send.device).cuda_stream
        if self._stream != stream:  # the same bookkeeping as use_torch_stream: later calls must see the change
            torch.cuda.synchronize(self.device)
            self.set_stream(stream)
            self._stream = stream
        self._check(self.lib.bkf_comm_allgather(self.h, C.c_void_p(send.data_ptr()), C.c_void_p(out.data_ptr()),
                                                send.numel(), code))
        return out

    def comm_destroy(self):
        self.lib.bkf_comm_destroy(self.h)
        self._comm_world = 0

    def close(self):
        sib = getattr(self, "_sibling", None)
        if sib is not None:
            sib.close()
            self._sibling = None
        if getattr(self, "h", None):
            self.lib.bkf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int | None):
        self.lib.bkf_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0))

    def use_torch_stream(self, bufs=None):
        """Device-pointer calls run on torch's current stream of this engine's device.  The workspace of the handle is
        shared by every call, so when the stream changes the old one is drained first; tensors of another device are
        refused (the kernels would run on the wrong GPU)."""
        import torch

        if bufs is not None and bufs.dev.index is not None and bufs.dev.index != self.device:
            raise BkfError(f"tensors live on cuda:{bufs.dev.index} but this engine drives cuda:{self.device}")
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            torch.cuda.synchronize(self.device)
            self.set_stream(s)
            self._stream = s

    def use_private_stream(self):
        """Give this engine a non-blocking CUDA stream of its own for host-pointer calls (kept alive by the engine): two
        engines with private streams overlap their copies and kernels (:meth:`logp_grad_overlapped`); the legacy default
        stream would serialise them."""
        import torch

        if getattr(self, "_own_stream", None) is None:
            self._own_stream = torch.cuda.Stream(device=self.device)
        return self._own_stream

    def use_default_stream(self):
        """Host-pointer calls run on the legacy default stream (a torch stream set earlier may be gone by now), or on the
        engine's private stream once :meth:`use_private_stream` was called."""
        own = getattr(self, "_own_stream", None)
        if own is not None:
            if self._stream != own.cuda_stream:
                import torch

                torch.cuda.synchronize(self.device)
                self.set_stream(own.cuda_stream)
                self._stream = own.cuda_stream
            return
        if self._stream != 0:
            import torch

            torch.cuda.synchronize(self.device)
            self.set_stream(0)
            self._stream = 0

    @property
    def launch_count(self) -> int:
        return int(self.lib.bkf_launch_count(self.h))

    @property
    def last_path(self) -> str:
        return self.lib.bkf_last_path(self.h).decode()

    def enable_timing(self, on=True):
        self.lib.bkf_enable_timing(self.h, int(on))
        self.lib.bkf_reset_timing(self.h)

    def kernel_times(self):
        """{kernel class: (total device ms, launches)} accumulated since enable_timing()."""
        out = {}
        for i, name in enumerate(("forward", "backward", "smoother")):
            ms, nl = C.c_double(0), C.c_longlong(0)
            self.lib.bkf_kernel_time_ms(self.h, i, C.byref(ms), C.byref(nl))
            out[name] = (ms.value, nl.value)
        return out

    def measure_fma_peak(self, dtype=F64) -> float:
        """Measured DFMA (F64) / FFMA (F32) pipe peak of this device in TFLOP/s."""
        v = C.c_double(0)
        self._check(self.lib.bkf_measure_fma_peak(self.h, int(dtype), C.byref(v)))
        return v.value

    def _check(self, rc):
        if rc != 0:
            raise BkfError(f"libbkf error {rc}: {self.lib.bkf_last_error(self.h).decode()}")

    # ------------------------------------------------------------------------------------------------
    def _problem(self, kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying=()):
        """Shape bookkeeping: returns (BkfProblem, y, [nine parameter arrays])."""
        k = KINDS[kind] if isinstance(kind, str) else int(kind)
        R = params["R"]
        Z = params["Z"]
        m, r = int(R.shape[-2]), int(R.shape[-1])
        p = int(Z.shape[-2])
        n = int(y.shape[-2])
        shared_mask, tv_mask = 0, 0
        arrs = []
        for i, name in enumerate(PARAM_NAMES):
            x = params[name]
            tv = name in time_varying
            if tv and name in ("a0", "P0"):
                raise BkfError(f"{name} is never time-varying")
            nd = _STATIC_NDIM[name] + int(tv)
            if tv:
                tv_mask |= 1 << i
            if x.ndim == nd:
                shared_mask |= 1 << i
            elif x.ndim == nd + 1:
                if x.shape[0] != B:
                    raise BkfError(f"{name}: leading batch dimension {x.shape[0]} != {B}")
            else:
                raise BkfError(f"{name}: expected {nd} (shared) or {nd + 1} (batched) dims, got {x.ndim}"
                               + ("" if tv else "; pass time_varying=(...) for matrices with a time axis"))
            if tv and x.shape[-_STATIC_NDIM[name] - 1] != n:
                raise BkfError(f"{name}: time axis has {x.shape[-_STATIC_NDIM[name] - 1]} steps, data has {n}")
            arrs.append(bufs.inp(x))
        y_shared = int(y.ndim == 2)
        if y.ndim == 3 and y.shape[0] != B:
            raise BkfError(f"y: leading batch dimension {y.shape[0]} != {B}")
        if y.shape[-1] != p:
            raise BkfError(f"y has {y.shape[-1]} observed series but Z has {p} rows")
        default_jitter = JITTER_F64 if bufs.code == F64 else JITTER_F32
        prob = BkfProblem(
            kind=k, dtype=bufs.code, mem=MEM_DEVICE if bufs.device else MEM_HOST, B=B, n=n, m=m, p=p, r=r,
            y_shared=y_shared, shared_mask=shared_mask, tv_mask=tv_mask,
            cov_jitter=default_jitter if cov_jitter is None else float(cov_jitter),
            missing_fill=MISSING_FILL if missing_fill_value is None else float(missing_fill_value),
        )
        return prob, bufs.inp(y), arrs

    @staticmethod
    def _batch_size(y, params, time_varying=()):
        B = None
        if y.ndim == 3:
            B = y.shape[0]
        for name in PARAM_NAMES:
            x = params[name]
            if x.ndim == _STATIC_NDIM[name] + 1 + int(name in time_varying):
                if B is not None and x.shape[0] != B:
                    raise BkfError(f"inconsistent batch sizes ({B} vs {x.shape[0]} for {name})")
                B = x.shape[0]
        return B

    # ------------------------------------------------------------------------------------------------
    def filter(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, cov_jitter=None, missing_fill_value=None,
               outputs=("filtered_states", "predicted_states", "observed_states", "filtered_covariances",
                        "predicted_covariances", "observed_covariances", "loglike_obs"), dtype=None,
               time_varying=()):
        """Batched ``BaseFilter.build_graph`` (kalman_filter.py:144-308).  Returns a dict of the requested
        outputs plus ``status``; unbatched inputs (no leading B anywhere) return unbatched outputs.

        ``time_varying``: names among c, d, T, Z, R, H, Q that carry a time axis of length n right after the optional
        batch axis (filters/utilities.py:8-47); step t then uses x[t] (standard and univariate filters)."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([y, *params.values()], dtype)
        params = {k_: (v if _is_torch(v) else np.asarray(v)) for k_, v in params.items()}
        y = y if _is_torch(y) else np.asarray(y)
        B = self._batch_size(y, params, time_varying)
        squeeze = B is None
        B = 1 if squeeze else int(B)
        prob, yb, arrs = self._problem(kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying)
        n, m, p = prob.n, prob.m, prob.p
        po = 1 if prob.kind == UNIVARIATE else p
        shapes = {
            "filtered_states": (B, n, m), "predicted_states": (B, n, m), "observed_states": (B, n, po),
            "filtered_covariances": (B, n, m, m), "predicted_covariances": (B, n, m, m),
            "observed_covariances": (B, n, po, po), "loglike_obs": (B, n),
        }
        outs = {k_: (bufs.out(shapes[k_]) if k_ in outputs else None) for k_ in shapes}
        status = bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_filter_forward(
            self.h, C.byref(prob), bufs.ptr(yb), *[bufs.ptr(a) for a in arrs],
            *[bufs.ptr(outs[k_]) for k_ in shapes], bufs.ptr(status))
        self._check(rc)
        res = {k_: v for k_, v in outs.items() if v is not None}
        res["status"] = status
        if squeeze:
            res = {k_: v[0] for k_, v in res.items()}
        return res

    @staticmethod
    def packed_layout(B, m, p, r):
        """Offsets (in elements) of [logp | g_a0 | g_P0 | g_c | g_d | g_T | g_Z | g_R | g_H | g_Q] inside one flat
        buffer of B * (1 + n_grad) elements: the single buffer the multi-GPU gather ships."""
        sizes = {"logp": 1, "a0": m, "P0": m * m, "c": m, "d": p, "T": m * m, "Z": p * m, "R": m * r,
                 "H": p * p, "Q": r * r}
        shapes = {"logp": (B,), "a0": (B, m), "P0": (B, m, m), "c": (B, m), "d": (B, p), "T": (B, m, m),
                  "Z": (B, p, m), "R": (B, m, r), "H": (B, p, p), "Q": (B, r, r)}
        off, layout = 0, {}
        for k_, sz in sizes.items():
            layout[k_] = (off, B * sz, shapes[k_])
            off += B * sz
        return layout, off

    def logp_grad(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, ll_cotangent=None, cov_jitter=None,
                  missing_fill_value=None, grads=PARAM_NAMES, dtype=None, packed_out=None, time_varying=(), outs=None):
        """logp[b] = sum_t loglike_obs[b,t] and d(sum_t cot*ll)/d{a0..Q} (reverse-mode adjoint kernels).

        ``packed_out``: optional flat preallocated buffer (see :meth:`packed_layout`); logp and the gradients
        are then written straight into it (views are returned), so a gather needs no packing kernel.
        ``outs``: optional dict of preallocated contiguous arrays ("logp", "status" and the gradient names) to write into
        (used by :meth:`logp_grad_overlapped`, whose chunks fill slices of one set of arrays)."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([y, *params.values()], dtype)
        params = {k_: (v if _is_torch(v) else np.asarray(v)) for k_, v in params.items()}
        y = y if _is_torch(y) else np.asarray(y)
        B = self._batch_size(y, params, time_varying)
        squeeze = B is None
        B = 1 if squeeze else int(B)
        prob, yb, arrs = self._problem(kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying)
        n, m, p, r = prob.n, prob.m, prob.p, prob.r
        if B > 1 and packed_out is None and grads:
            need = C.c_size_t(0)
            self._check(self.lib.bkf_logp_grad_workspace_bytes(C.byref(prob), C.byref(need)))
            budget = self.max_workspace_bytes
            if budget is None:
                budget = int(0.7 * self.device_memory()[1])
            if need.value > budget:
                return self._logp_grad_in_slices(kind, y, params, ll_cotangent, B, -(-need.value // budget),
                                                 dict(cov_jitter=cov_jitter, missing_fill_value=missing_fill_value,
                                                      grads=grads, dtype=dtype, time_varying=time_varying))
        gshape = {"a0": (B, m), "P0": (B, m, m), "c": (B, m), "d": (B, p), "T": (B, m, m), "Z": (B, p, m),
                  "R": (B, m, r), "H": (B, p, p), "Q": (B, r, r)}
        for k_ in time_varying:  # gradients of time-varying parameters keep the time axis
            gshape[k_] = (B, n) + gshape[k_][1:]
        if packed_out is not None and time_varying:
            raise BkfError("packed_out is laid out for time-invariant parameters")
        if packed_out is not None:
            layout, total = self.packed_layout(B, m, p, r)
            flat = packed_out.reshape(-1)
            if flat.shape[0] != total:
                raise BkfError(f"packed_out must hold {total} elements, got {flat.shape[0]}")
            view = lambda k_: flat[layout[k_][0]:layout[k_][0] + layout[k_][1]].reshape(layout[k_][2])
            gouts = {k_: (view(k_) if k_ in grads else None) for k_ in PARAM_NAMES}
            logp = view("logp")
        elif outs is not None:
            for k_ in ("logp", "status", *grads):
                want = {"logp": (B,), "status": (B,)}.get(k_) or gshape[k_]
                if tuple(outs[k_].shape) != tuple(want) or not outs[k_].flags["C_CONTIGUOUS"]:
                    raise BkfError(f"outs[{k_!r}] must be a contiguous array of shape {want}")
            gouts = {k_: (outs[k_] if k_ in grads else None) for k_ in PARAM_NAMES}
            logp = outs["logp"]
        else:
            gouts = {k_: (bufs.out(gshape[k_]) if k_ in grads else None) for k_ in PARAM_NAMES}
            logp = bufs.out((B,))
        status = outs["status"] if outs is not None and packed_out is None else bufs.out_int((B,))
        cot = None
        if ll_cotangent is not None:
            cot = bufs.inp(ll_cotangent)
            if tuple(cot.shape) != (B, n) and not (squeeze and tuple(cot.shape) == (n,)):
                raise BkfError("ll_cotangent must have shape (B, n)")
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_filter_logp_grad(
            self.h, C.byref(prob), bufs.ptr(yb), *[bufs.ptr(a) for a in arrs], bufs.ptr(cot), bufs.ptr(logp),
            *[bufs.ptr(gouts[k_]) for k_ in PARAM_NAMES], bufs.ptr(status))
        self._check(rc)
        g = {k_: v for k_, v in gouts.items() if v is not None}
        if squeeze:
            return logp[0], {k_: v[0] for k_, v in g.items()}, status[0]
        return logp, g, status

    def logp_grad_overlapped(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, chunks=2, ll_cotangent=None, cov_jitter=None,
                             missing_fill_value=None, dtype=None):
        """:meth:`logp_grad` for HOST arrays with the host <-> device copies hidden: the batch goes through in ``chunks``
        contiguous slices, alternately on this engine and on a sibling engine (second handle, second workspace), each on a
        private non-blocking stream and driven by its own host thread (ctypes releases the GIL), so the H2D copy of one
        slice and the D2H copy of another run under the kernels of a third.  Series are independent: the results are
        bit-identical to one call.  Inputs should be page-locked (``bkf_host_alloc`` / ``torch`` pinned memory) for the
        copies to be asynchronous; the outputs are."""
        import threading

        params = {k_: np.asarray(v) for k_, v in zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q))}
        y = np.asarray(y)
        if any(_is_torch(v) for v in (y, *params.values())):
            raise BkfError("logp_grad_overlapped takes host (NumPy) arrays")
        B = self._batch_size(y, params, ())
        chunks = int(chunks)
        if B is None or chunks < 2 or int(B) < 2 * chunks:
            return self.logp_grad(kind, y, *params.values(), ll_cotangent=ll_cotangent, cov_jitter=cov_jitter,
                                  missing_fill_value=missing_fill_value, dtype=dtype)
        B = int(B)
        nd = np.dtype(np.float64 if dtype is None else dtype)
        m, r = params["R"].shape[-2:]
        p = params["Z"].shape[-2]
        pool = pinned_pool()
        shapes = {"a0": (B, m), "P0": (B, m, m), "c": (B, m), "d": (B, p), "T": (B, m, m), "Z": (B, p, m), "R": (B, m, r),
                  "H": (B, p, p), "Q": (B, r, r)}
        outs = {k_: pool.empty(shapes[k_], nd) for k_ in PARAM_NAMES}
        outs["logp"] = pool.empty((B,), nd)
        outs["status"] = np.zeros((B,), dtype=np.int32)
        if getattr(self, "_sibling", None) is None:
            self._sibling = KalmanEngine(self.device)
        engines = (self, self._sibling)
        for e in engines:
            e.use_private_stream()
        bounds = [B * i // chunks // 32 * 32 for i in range(chunks)] + [B]
        batched = lambda name, x: x.ndim == _STATIC_NDIM[name] + 1
        errors = []

        def work(which):
            try:
                for i in range(which, chunks, 2):
                    sl = slice(bounds[i], bounds[i + 1])
                    args = [params[k_][sl] if batched(k_, params[k_]) else params[k_] for k_ in PARAM_NAMES]
                    engines[which].logp_grad(
                        kind, y[sl] if y.ndim == 3 else y, *args,
                        ll_cotangent=None if ll_cotangent is None else np.asarray(ll_cotangent)[sl], cov_jitter=cov_jitter,
                        missing_fill_value=missing_fill_value, dtype=dtype, outs={k_: v[sl] for k_, v in outs.items()})
            except Exception as exc:  # re-raised in the caller's thread
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(w,)) for w in (0, 1)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        return outs["logp"], {k_: outs[k_] for k_ in PARAM_NAMES}, outs["status"]

    def forward_saved(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, cov_jitter=None, missing_fill_value=None, dtype=None,
                      time_varying=(), want_ll=True):
        """Forward sweep only (``bkf_filter_forward_saved``): returns ``(loglike_obs or None, logp, status, ticket)`` and
        keeps the sweep's record on the device for ONE later :meth:`backward_saved` with that ticket."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([y, *params.values()], dtype)
        params = {k_: (v if _is_torch(v) else np.asarray(v)) for k_, v in params.items()}
        y = y if _is_torch(y) else np.asarray(y)
        B = self._batch_size(y, params, time_varying)
        squeeze = B is None
        B = 1 if squeeze else int(B)
        prob, yb, arrs = self._problem(kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying)
        ll = bufs.out((B, prob.n)) if want_ll else None
        logp = bufs.out((B,))
        status = bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        ticket = C.c_ulonglong(0)
        self._check(self.lib.bkf_filter_forward_saved(
            self.h, C.byref(prob), bufs.ptr(yb), *[bufs.ptr(a) for a in arrs], bufs.ptr(ll), bufs.ptr(logp),
            bufs.ptr(status), C.byref(ticket)))
        gshape = {"a0": (B, prob.m), "P0": (B, prob.m, prob.m), "c": (B, prob.m), "d": (B, prob.p),
                  "T": (B, prob.m, prob.m), "Z": (B, prob.p, prob.m), "R": (B, prob.m, prob.r),
                  "H": (B, prob.p, prob.p), "Q": (B, prob.r, prob.r)}
        for k_ in time_varying:
            gshape[k_] = (B, prob.n) + gshape[k_][1:]
        self._saved = (ticket.value, bufs, gshape, squeeze, (B, prob.n))
        if squeeze:
            return (None if ll is None else ll[0]), logp[0], status[0], ticket.value
        return ll, logp, status, ticket.value

    def backward_saved(self, ticket, ll_cotangent=None, grads=PARAM_NAMES):
        """Adjoint sweep on the record of :meth:`forward_saved` (``bkf_filter_backward_saved``).  Returns the dict of
        gradients, or ``None`` when the record is gone (the caller then runs :meth:`logp_grad`)."""
        saved, self._saved = self._saved, None
        if saved is None or saved[0] != int(ticket):
            return None
        _, bufs, gshape, squeeze, (B, n) = saved
        gouts = {k_: (bufs.out(gshape[k_]) if k_ in grads else None) for k_ in PARAM_NAMES}
        cot = None
        if ll_cotangent is not None:
            cot = bufs.inp(ll_cotangent)
            if tuple(cot.shape) != (B, n) and not (squeeze and tuple(cot.shape) == (n,)):
                raise BkfError("ll_cotangent must have shape (B, n)")
        rc = self.lib.bkf_filter_backward_saved(self.h, C.c_ulonglong(int(ticket)), bufs.ptr(cot),
                                                *[bufs.ptr(gouts[k_]) for k_ in PARAM_NAMES])
        if rc == ERR_STALE:
            return None
        self._check(rc)
        g = {k_: v for k_, v in gouts.items() if v is not None}
        return {k_: v[0] for k_, v in g.items()} if squeeze else g

    def logp_grad_compact(self, kind, y, base, positions, u, ll_cotangent=None, cov_jitter=None,
                          missing_fill_value=None, dtype=None, want_grad=True):
        """logp and gradient for B draws that differ only in a few matrix entries.

        ``base``: dict of the nine matrices without batch axis; ``positions``: list of ``(name, index_tuple)`` -- entry k
        of ``u[b]`` is written into ``base[name][index_tuple]`` for draw b; ``u``: (B, len(positions)).  Returns
        ``(logp (B,), g_u (B, nnz), status (B,))`` where ``g_u[b, k] = d logp_b / d u[b, k]``.  Only ``u``, ``logp`` and
        ``g_u`` cross the host-device boundary (``bkf_filter_logp_grad_compact``)."""
        bufs = _Buffers([y, u, *base.values()], dtype)
        conv = lambda v: v if _is_torch(v) else np.asarray(v)
        y, u = conv(y), conv(u)
        base = {k_: conv(base[k_]) for k_ in PARAM_NAMES}
        for k_ in PARAM_NAMES:
            if base[k_].ndim != _STATIC_NDIM[k_]:
                raise BkfError(f"base[{k_!r}] must not have a batch axis")
        B, nnz = int(u.shape[0]), int(u.shape[1])
        if nnz != len(positions):
            raise BkfError("u must have one column per position")
        prob, yb, arrs = self._problem(kind, bufs, y, base, B, cov_jitter, missing_fill_value)
        which = (C.c_int * max(nnz, 1))()
        index = (C.c_int * max(nnz, 1))()
        for k, (name, idx) in enumerate(positions):
            which[k] = PARAM_NAMES.index(name)
            index[k] = int(np.ravel_multi_index(tuple(np.atleast_1d(idx)), base[name].shape))
        ub = bufs.inp(u)
        cot = None if ll_cotangent is None else bufs.inp(ll_cotangent)
        logp = bufs.out((B,))
        g_u = bufs.out((B, nnz)) if want_grad else None
        status = bufs.out_int((B,))
        base_ptrs = (C.c_void_p * 9)(*[bufs.ptr(a).value for a in arrs])
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_filter_logp_grad_compact(
            self.h, C.byref(prob), bufs.ptr(yb), base_ptrs, nnz, which, index, bufs.ptr(ub), bufs.ptr(cot),
            bufs.ptr(logp), bufs.ptr(g_u), bufs.ptr(status))
        self._check(rc)
        return logp, g_u, status

    def logp_grad_theta(self, kind, y, base, positions, terms, theta, ll_cotangent=None, cov_jitter=None,
                        missing_fill_value=None, dtype=None, want_grad=True):
        """logp and its gradient with respect to a model's FREE PARAMETERS for B draws (the rest of SURVEY section 8 row
        f-1): ``theta`` (B, ntheta) is all that goes in per draw, the matrices are assembled on the device.

        ``base`` / ``positions`` as in ``logp_grad_compact``; ``terms``: list of ``(k, coef, [(theta_index, fn), ...])``
        -- entry k of the compact vector is the sum of its terms ``coef * prod fn(theta[theta_index])`` (at most three
        factors; ``fn`` one of ``FN_ID, FN_COS, FN_SIN, FN_EXP, FN_ONE_MINUS, FN_SQUARE``).  The builders of
        ``pymc_experimental_b200.theta_maps`` produce these for the reference's model families.  Returns
        ``(logp (B,), g_theta (B, ntheta), status (B,))`` (``bkf_filter_logp_grad_theta``)."""
        bufs = _Buffers([y, theta, *base.values()], dtype)
        conv = lambda v: v if _is_torch(v) else np.asarray(v)
        y, theta = conv(y), conv(theta)
        base = {k_: conv(base[k_]) for k_ in PARAM_NAMES}
        for k_ in PARAM_NAMES:
            if base[k_].ndim != _STATIC_NDIM[k_]:
                raise BkfError(f"base[{k_!r}] must not have a batch axis")
        B, ntheta = int(theta.shape[0]), int(theta.shape[1])
        nnz, nterms = len(positions), len(terms)
        prob, yb, arrs = self._problem(kind, bufs, y, base, B, cov_jitter, missing_fill_value)
        which = (C.c_int * max(nnz, 1))()
        index = (C.c_int * max(nnz, 1))()
        for k, (name, idx) in enumerate(positions):
            which[k] = PARAM_NAMES.index(name)
            index[k] = int(np.ravel_multi_index(tuple(np.atleast_1d(idx)), base[name].shape))
        t_entry = (C.c_int * max(nterms, 1))()
        t_coef = (C.c_double * max(nterms, 1))()
        t_nfac = (C.c_int * max(nterms, 1))()
        t_theta = (C.c_int * max(3 * nterms, 1))()
        t_fn = (C.c_int * max(3 * nterms, 1))()
        for q, (k, coef, factors) in enumerate(terms):
            if len(factors) > 3:
                raise BkfError("a term has at most three factors")
            t_entry[q], t_coef[q], t_nfac[q] = int(k), float(coef), len(factors)
            for f, (ti, fn) in enumerate(factors):
                t_theta[3 * q + f], t_fn[3 * q + f] = int(ti), int(fn)
        tb = bufs.inp(theta)
        cot = None if ll_cotangent is None else bufs.inp(ll_cotangent)
        logp = bufs.out((B,))
        g_theta = bufs.out((B, ntheta)) if want_grad else None
        status = bufs.out_int((B,))
        base_ptrs = (C.c_void_p * 9)(*[bufs.ptr(a).value for a in arrs])
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_filter_logp_grad_theta(
            self.h, C.byref(prob), bufs.ptr(yb), base_ptrs, nnz, which, index, nterms, t_entry, t_coef, t_nfac, t_theta,
            t_fn, ntheta, bufs.ptr(tb), bufs.ptr(cot), bufs.ptr(logp), bufs.ptr(g_theta), bufs.ptr(status))
        self._check(rc)
        return logp, g_theta, status

    # ------------------------------------------------------------------------------------------------
    def _stationary_problem(self, bufs, c, T, R, Q):
        conv = lambda v: v if _is_torch(v) else np.asarray(v)
        vals = {"c": conv(c), "T": conv(T), "R": conv(R), "Q": conv(Q)}
        B = None
        for name, x in vals.items():
            if x.ndim == _STATIC_NDIM[name] + 1:
                if B is not None and x.shape[0] != B:
                    raise BkfError(f"inconsistent batch sizes ({B} vs {x.shape[0]} for {name})")
                B = int(x.shape[0])
            elif x.ndim != _STATIC_NDIM[name]:
                raise BkfError(f"{name}: expected {_STATIC_NDIM[name]} or {_STATIC_NDIM[name] + 1} dims")
        squeeze = B is None
        B = 1 if squeeze else B
        m, r = int(vals["R"].shape[-2]), int(vals["R"].shape[-1])
        shared = 0
        for name in ("c", "T", "R", "Q"):
            if vals[name].ndim == _STATIC_NDIM[name]:
                shared |= 1 << PARAM_NAMES.index(name)
        prob = BkfProblem(kind=STANDARD, dtype=bufs.code, mem=MEM_DEVICE if bufs.device else MEM_HOST, B=B, n=1, m=m,
                          p=1, r=r, y_shared=0, shared_mask=shared, tv_mask=0, cov_jitter=0.0, missing_fill=MISSING_FILL)
        return prob, {k_: bufs.inp(v) for k_, v in vals.items()}, B, m, r, squeeze

    def stationary_init(self, c, T, R, Q, dtype=None):
        """``a0 = (I - T)^-1 c`` and ``P0 = solve_discrete_lyapunov(T, R Q R^T)`` for a batch of models
        (models/SARIMAX.py:369-381, VARMAX.py:379-393).  Returns ``(a0, P0, status)``."""
        bufs = _Buffers([c, T, R, Q], dtype)
        prob, v, B, m, r, squeeze = self._stationary_problem(bufs, c, T, R, Q)
        a0, P0, status = bufs.out((B, m)), bufs.out((B, m, m)), bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        self._check(self.lib.bkf_stationary_init(self.h, C.byref(prob), *[bufs.ptr(v[k_]) for k_ in ("c", "T", "R", "Q")],
                                                 bufs.ptr(a0), bufs.ptr(P0), bufs.ptr(status)))
        return (a0[0], P0[0], status[0]) if squeeze else (a0, P0, status)

    def stationary_init_grad(self, T, R, Q, a0, P0, a0_bar, P0_bar, dtype=None):
        """Pull-back of :meth:`stationary_init`: ``{c, T, R, Q: gradient}`` of ``<a0_bar, a0> + <P0_bar, P0>``."""
        bufs = _Buffers([T, R, Q, a0, P0], dtype)
        prob, v, B, m, r, squeeze = self._stationary_problem(bufs, np.zeros(int(np.shape(T)[-1])), T, R, Q)
        fix = lambda x, nd: None if x is None else bufs.inp(x if not squeeze else x[None])
        a0b, P0b, ab, Pb = fix(a0, 1), fix(P0, 2), fix(a0_bar, 1), fix(P0_bar, 2)
        g = {"c": bufs.out((B, m)), "T": bufs.out((B, m, m)), "R": bufs.out((B, m, r)), "Q": bufs.out((B, r, r))}
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        self._check(self.lib.bkf_stationary_init_grad(
            self.h, C.byref(prob), bufs.ptr(v["T"]), bufs.ptr(v["R"]), bufs.ptr(v["Q"]), bufs.ptr(a0b), bufs.ptr(P0b),
            bufs.ptr(ab), bufs.ptr(Pb), *[bufs.ptr(g[k_]) for k_ in ("c", "T", "R", "Q")]))
        return {k_: x[0] for k_, x in g.items()} if squeeze else g

    @staticmethod
    def _rng(prob, seed, offset):
        if seed is None:
            raise BkfError("pass standard normal variates z, or seed= (and offset=) for the in-kernel generator")
        prob.rng_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        prob.rng_offset = int(offset) & 0xFFFFFFFFFFFFFFFF

    def sample_mvn(self, mu, cov, z=None, dtype=None, seed=None, offset=0):
        """One draw from ``N(mu[..., t, :], cov[..., t, :, :])`` per (series, step): ``x = mu + L z`` with ``L L^T = cov``
        (``bkf_mvn_sample``; SURVEY section 8 row f-3 -- the batched form of ``SequenceMvNormal.rv_op`` /
        ``MvNormalSVD``, filters/distributions.py:52-87, 411-443).  ``mu, z``: (B, n, m) or (n, m); ``cov``: (B, n, m, m)
        or (n, m, m); ``z`` standard normal variates, or ``None``: drawn in the kernel by the counter-based generator
        (Philox4x32-10 + Box-Muller) keyed by ``seed`` at counter ``offset`` -- no variates cross PCIe, and variate i is
        the same whatever the batch size or launch geometry.  Returns ``(x, status)``."""
        bufs = _Buffers([mu, cov, z], dtype)
        conv = lambda v: v if (v is None or _is_torch(v)) else np.asarray(v)
        mu, cov, z = conv(mu), conv(cov), conv(z)
        squeeze = mu.ndim == 2
        if squeeze:
            mu, cov, z = mu[None], cov[None], (None if z is None else z[None])
        if mu.ndim != 3 or cov.ndim != 4 or (z is not None and tuple(z.shape) != tuple(mu.shape)) or \
                tuple(cov.shape) != tuple(mu.shape) + (mu.shape[-1],):
            raise BkfError("sample_mvn: expected mu, z (B, n, m) and cov (B, n, m, m)")
        B, n, m = (int(v) for v in mu.shape)
        prob = BkfProblem(kind=STANDARD, dtype=bufs.code, mem=MEM_DEVICE if bufs.device else MEM_HOST, B=B, n=n, m=m,
                          p=1, r=1, y_shared=0, shared_mask=0, tv_mask=0, cov_jitter=0.0, missing_fill=MISSING_FILL)
        dmu, dcov, dz = bufs.inp(mu), bufs.inp(cov), (None if z is None else bufs.inp(z))
        if z is None:
            self._rng(prob, seed, offset)
        x, status = bufs.out((B, n, m)), bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        self._check(self.lib.bkf_mvn_sample(self.h, C.byref(prob), bufs.ptr(dmu), bufs.ptr(dcov), bufs.ptr(dz),
                                            bufs.ptr(x), bufs.ptr(status)))
        return (x[0], status[0]) if squeeze else (x, status)

    def sample_conditional_posterior(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, z=None, cov_jitter=None,
                                     missing_fill_value=None, dtype=None, time_varying=(), seed=None, offset=0):
        """Filter, smoother and one draw of the hidden states per (series, step) from the smoothed moments in ONE call
        (``bkf_filter_smooth_sample``; ``sample_conditional_posterior``, core/statespace.py:1061-1171): the smoothed
        covariances stay on the device.  ``z``: standard normal variates (B, n, m), or ``None`` with ``seed`` / ``offset``:
        drawn in the kernel (see ``sample_mvn``).  Returns ``(x, loglike_obs, status)``."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([y, z, *params.values()], dtype)
        params = {k_: (v if _is_torch(v) else np.asarray(v)) for k_, v in params.items()}
        y = y if _is_torch(y) else np.asarray(y)
        z = z if (z is None or _is_torch(z)) else np.asarray(z)
        B = self._batch_size(y, params, time_varying)
        squeeze = B is None
        B = 1 if squeeze else int(B)
        prob, yb, arrs = self._problem(kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying)
        n, m = prob.n, prob.m
        if z is None:
            self._rng(prob, seed, offset)
            dz = None
        else:
            if squeeze:
                z = z[None]
            if tuple(z.shape) != (B, n, m):
                raise BkfError(f"z: expected shape {(B, n, m)}, got {tuple(z.shape)}")
            dz = bufs.inp(z)
        x, ll, status = bufs.out((B, n, m)), bufs.out((B, n)), bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        self._check(self.lib.bkf_filter_smooth_sample(self.h, C.byref(prob), bufs.ptr(yb), *[bufs.ptr(a) for a in arrs],
                                                      bufs.ptr(dz), bufs.ptr(ll), bufs.ptr(x), bufs.ptr(status)))
        return (x[0], ll[0], status[0]) if squeeze else (x, ll, status)

    def simulate(self, a0, P0, c, d, T, Z, R, H, Q, z_x0=None, z_y0=None, z_state=None, z_obs=None, dtype=None,
                 time_varying=(), batch=None, steps=None, seed=None, offset=0):
        """Unconditional simulation of B models (``bkf_simulate``; ``LinearGaussianStateSpace.rv_op``,
        filters/distributions.py:181-293).  ``z_x0`` (B, m), ``z_y0`` (B, p), ``z_state`` (B, n, r), ``z_obs`` (B, n, p):
        standard normal variates (they define B and the number of steps n).  Parameters may be batched or shared;
        ``time_varying`` names those with a time axis of length n (the reference's ``sequence_names``).
        With no variates, ``batch`` series of ``steps`` steps are drawn by the in-kernel generator (``seed`` / ``offset``,
        see ``sample_mvn``).  Returns ``(states (B, n+1, m), obs (B, n+1, p), status)``; row 0 is the initial draw."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([z_x0, z_y0, z_state, z_obs, *params.values()], dtype)
        conv = lambda v: v if _is_torch(v) else np.asarray(v)
        params = {k_: conv(v) for k_, v in params.items()}
        given = [v is not None for v in (z_x0, z_y0, z_state, z_obs)]
        if any(given) and not all(given):
            raise BkfError("simulate: pass all four variate arrays, or none (in-kernel generator)")
        if all(given):
            z_x0, z_y0, z_state, z_obs = conv(z_x0), conv(z_y0), conv(z_state), conv(z_obs)
            if z_state.ndim != 3 or z_obs.ndim != 3 or z_x0.ndim != 2 or z_y0.ndim != 2:
                raise BkfError("simulate: expected z_x0 (B, m), z_y0 (B, p), z_state (B, n, r), z_obs (B, n, p)")
            B, n = int(z_state.shape[0]), int(z_state.shape[1])
            shape_y = z_obs
        else:
            if batch is None or steps is None:
                raise BkfError("simulate: without variates, pass batch= and steps=")
            B, n = int(batch), int(steps)
            shape_y = np.broadcast_to(np.zeros((), dtype=np.float32), (B, n, int(params["Z"].shape[-2])))
        prob, _, arrs = self._problem(STANDARD, bufs, shape_y, params, B, None, None, time_varying)
        m, p, r = prob.m, prob.p, prob.r
        if all(given):
            if tuple(z_x0.shape) != (B, m) or tuple(z_y0.shape) != (B, p) or tuple(z_state.shape) != (B, n, r) or \
                    tuple(z_obs.shape) != (B, n, p):
                raise BkfError("simulate: variate shapes do not match the model (m, p, r) / batch / steps")
            zs = [bufs.inp(v) for v in (z_x0, z_y0, z_state, z_obs)]
        else:
            self._rng(prob, seed, offset)
            zs = [None] * 4
        states, obs, status = bufs.out((B, n + 1, m)), bufs.out((B, n + 1, p)), bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        self._check(self.lib.bkf_simulate(self.h, C.byref(prob), *[bufs.ptr(x) for x in arrs], *[bufs.ptr(x) for x in zs],
                                          bufs.ptr(states), bufs.ptr(obs), bufs.ptr(status)))
        return states, obs, status

    def _logp_grad_in_slices(self, kind, y, params, ll_cotangent, B, nslices, kw):
        """Series are independent: a batch whose trajectory workspace would not fit the device is processed in
        contiguous slices of the batch axis (shared inputs are passed to every slice) and the results concatenated."""
        tv = kw["time_varying"]
        per = -(-B // int(nslices))
        parts = []
        for b0 in range(0, B, per):
            sl = slice(b0, min(B, b0 + per))
            batched = lambda name, x: x.ndim == _STATIC_NDIM[name] + 1 + int(name in tv)
            args = [params[k_][sl] if batched(k_, params[k_]) else params[k_] for k_ in PARAM_NAMES]
            parts.append(self.logp_grad(kind, y[sl] if y.ndim == 3 else y, *args,
                                        ll_cotangent=None if ll_cotangent is None else ll_cotangent[sl], **kw))
        if _is_torch(parts[0][0]):
            import torch

            cat = torch.cat
        else:
            cat = np.concatenate
        logp = cat([q[0] for q in parts])
        grads = {k_: cat([q[1][k_] for q in parts]) for k_ in parts[0][1]}
        status = cat([q[2] for q in parts])
        return logp, grads, status

    def smooth(self, T, R, Q, filtered_states, filtered_covariances, cov_jitter=None, dtype=None, time_varying=(),
               return_status=False):
        """Batched ``KalmanSmoother.build_graph`` (kalman_smoother.py:66-105).  Time-varying T / R / Q
        ([B,] n, ...) follow the reference's pairing filtered[t] <-> x[t+1].
        ``return_status``: also return the per-series status word (``STATUS_SMOOTHER_ILL``: the series was redone with
        the eigen-decomposition pinv)."""
        bufs = _Buffers([T, R, Q, filtered_states, filtered_covariances], dtype)
        conv = lambda v: v if _is_torch(v) else np.asarray(v)
        T, R, Q, fa, fP = map(conv, (T, R, Q, filtered_states, filtered_covariances))
        squeeze = fa.ndim == 2
        if squeeze:
            fa, fP = fa[None], fP[None]
        B, n, m = (int(s) for s in fa.shape)
        r = int(R.shape[-1])
        shared_mask, tv_mask = 0, 0
        for name, x, bit in (("T", T, 4), ("R", R, 6), ("Q", Q, 8)):
            tv = name in time_varying
            if tv:
                tv_mask |= 1 << bit
                if x.shape[-3] != n:
                    raise BkfError(f"{name}: time axis has {x.shape[-3]} steps, the filtered trajectory has {n}")
            if x.ndim == 2 + int(tv):
                shared_mask |= 1 << bit
            elif x.shape[0] != B:
                raise BkfError(f"{name}: leading batch dimension {x.shape[0]} != {B}")
        default_jitter = JITTER_F64 if bufs.code == F64 else JITTER_F32
        prob = BkfProblem(kind=STANDARD, dtype=bufs.code, mem=MEM_DEVICE if bufs.device else MEM_HOST, B=B, n=n,
                          m=m, p=1, r=r, y_shared=0, shared_mask=shared_mask, tv_mask=tv_mask,
                          cov_jitter=default_jitter if cov_jitter is None else float(cov_jitter),
                          missing_fill=MISSING_FILL)
        Tb, Rb, Qb, fab, fPb = map(bufs.inp, (T, R, Q, fa, fP))
        sa, sP = bufs.out((B, n, m)), bufs.out((B, n, m, m))
        status = bufs.out_int((B,)) if return_status else None
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_smooth(self.h, C.byref(prob), *[bufs.ptr(x) for x in (Tb, Rb, Qb, fab, fPb, sa, sP)],
                                 bufs.ptr(status) if return_status else None)
        self._check(rc)
        out = (sa[0], sP[0]) if squeeze else (sa, sP)
        if return_status:
            out = out + ((status[0] if squeeze else status),)
        return out

    def filter_smooth(self, kind, y, a0, P0, c, d, T, Z, R, H, Q, cov_jitter=None, missing_fill_value=None,
                      dtype=None, time_varying=()):
        """Filter forward + RTS smoother in one call (core/statespace.py:903-951); the filtered trajectory
        stays in the device workspace.  Returns (loglike_obs, smoothed_states, smoothed_covariances, status)."""
        params = dict(zip(PARAM_NAMES, (a0, P0, c, d, T, Z, R, H, Q)))
        bufs = _Buffers([y, *params.values()], dtype)
        params = {k_: (v if _is_torch(v) else np.asarray(v)) for k_, v in params.items()}
        y = y if _is_torch(y) else np.asarray(y)
        B = self._batch_size(y, params, time_varying)
        squeeze = B is None
        B = 1 if squeeze else int(B)
        prob, yb, arrs = self._problem(kind, bufs, y, params, B, cov_jitter, missing_fill_value, time_varying)
        n, m = prob.n, prob.m
        ll, sa, sP = bufs.out((B, n)), bufs.out((B, n, m)), bufs.out((B, n, m, m))
        status = bufs.out_int((B,))
        if bufs.device:
            self.use_torch_stream(bufs)
        else:
            self.use_default_stream()
        rc = self.lib.bkf_filter_smooth(self.h, C.byref(prob), bufs.ptr(yb), *[bufs.ptr(a) for a in arrs],
                                        bufs.ptr(ll), bufs.ptr(sa), bufs.ptr(sP), bufs.ptr(status))
        self._check(rc)
        if squeeze:
            return ll[0], sa[0], sP[0], status[0]
        return ll, sa, sP, status


_default_engines: dict[int, KalmanEngine] = {}
_cuda_pid: int | None = None  # the process that created the first engine (and with it a CUDA context)


def _after_fork_in_child():
    """A forked child inherits handles and pinned buffers that belong to the parent's CUDA context: drop them (without
    calling into the driver) so that nothing dangling is ever used."""
    global _pinned_pool
    for eng in _default_engines.values():
        eng.h = None
    _default_engines.clear()
    _pinned_pool = None


if hasattr(os, "register_at_fork"):
    os.register_at_fork(after_in_child=_after_fork_in_child)


def default_engine(device: int = 0) -> KalmanEngine:
    """Lazily created per-process engine, as a PyTensor ``Op.perform`` needs it.

    NOT fork-safe, and says so: CUDA cannot be initialised in a child forked from a process that already holds a CUDA
    context, and ``pm.sample`` evaluates the model in the parent (start-value check) before it forks its chain
    workers.  Use ``pm.sample(..., mp_ctx="spawn")`` (or ``"forkserver"``, or ``cores=1``): a spawned worker imports this
    module afresh and gets an engine of its own here."""
    global _cuda_pid
    pid = os.getpid()
    if _cuda_pid is not None and _cuda_pid != pid:
        raise BkfError(
            f"this process (pid {pid}) was forked from pid {_cuda_pid}, which had already initialised CUDA: the Kalman "
            "engine cannot run here.  Start the workers with the 'spawn' or 'forkserver' context "
            "(pm.sample(..., mp_ctx='spawn')) or sample with cores=1")
    if device not in _default_engines:
        _default_engines[device] = KalmanEngine(device)
        _cuda_pid = pid
    return _default_engines[device]
