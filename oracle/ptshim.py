"""Eager PyTensor-API shim: runs the reference's OWN filter source files on torch.float64 tensors.

TEST INFRASTRUCTURE ONLY (see oracle/kalman_np.py header).  PyTensor / PyMC are not installable in this
environment, so ``pymc_extras.statespace.filters`` cannot be imported as-is.  This module
registers a *minimal eager stand-in* for the subset of the PyTensor API that the four hot-path source files use

    pymc_extras/statespace/utils/constants.py
    pymc_extras/statespace/filters/utilities.py
    pymc_extras/statespace/filters/kalman_filter.py
    pymc_extras/statespace/filters/kalman_smoother.py

and then executes those files from where they lie under ``/root/reference`` (nothing is copied).  Every
"symbolic" variable is an eager ``Var`` wrapping a ``torch.float64`` tensor, ``pytensor.scan`` is a Python
loop, and linalg Ops map to the LAPACK-backed torch/numpy routines.  Because the tensors are torch tensors,
``torch.autograd`` differentiates straight through the reference's own code, giving the gradient golden
vectors as well.

Used only by ``tests/golden/generate_golden.py`` (in the build container, where /root/reference exists) to
produce the committed fixtures in ``tests/golden/*.npz``.  It never runs on the GPU box.

Fidelity notes (what the stand-in assumes about PyTensor):
* ``pt.linalg.solve(assume_a="pos")``: forward reads the upper triangle like LAPACK ``?posv(uplo='U')``
  (PyTensor ``Solve.perform`` -> ``scipy.linalg.solve(lower=False)``); backward is PyTensor's generic
  ``SolveBase.L_op`` (``b_bar = solve(A^T, c_bar)``, ``A_bar = -outer(b_bar, c)``, no symmetrisation).
* ``pt.linalg.det``: ``Det.grad`` = ``g * det * inv(A)^T``  (== torch's).
* ``pt.linalg.cholesky``: ``Cholesky.L_op`` returns ``tril(s + s.T) - diag(s)`` (gradient in the lower
  triangle only); reproduced by differentiating through an explicit lower->symmetric fill.
* ``pt.linalg.qr(mode="r")``: LAPACK Householder R; differentiated through torch's reduced QR.
* ``pt.linalg.pinv``: ``numpy.linalg.pinv`` (what ``PInv.perform`` calls); not differentiated.
* ``ifelse`` / ``switch``: eager Python selection / ``torch.where``.
"""

from __future__ import annotations

import importlib.util
import sys
import types

import numpy as np
import torch

DT = torch.float64


def _raw(x):
    if isinstance(x, Var):
        return x.v
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, float, np.floating, np.integer)):
        return torch.tensor(float(x), dtype=DT)
    return torch.as_tensor(np.asarray(x))


class _Type:
    def __init__(self, v):
        self.shape = tuple(v.shape)
        self.dtype = "float64" if v.dtype == DT else str(v.dtype).replace("torch.", "")
        self.ndim = v.ndim


class Var:
    """Eager stand-in for pytensor.tensor.TensorVariable."""

    def __init__(self, v, sub=None):
        self.v = v
        self.name = None
        self._sub = sub  # (base Var, index) when produced by __getitem__, for set_subtensor

    # --- metadata ---
    @property
    def type(self):
        return _Type(self.v)

    @property
    def shape(self):
        return tuple(self.v.shape)

    @property
    def ndim(self):
        return self.v.ndim

    @property
    def dtype(self):
        return self.type.dtype

    @property
    def T(self):
        return Var(self.v.T if self.v.ndim == 2 else self.v)

    # --- arithmetic ---
    def _bin(self, other, fn):
        a, b = self.v, _raw(other)
        if a.dtype == torch.bool and b.dtype != torch.bool:
            a = a.to(b.dtype)
        if b.dtype == torch.bool and a.dtype != torch.bool:
            b = b.to(a.dtype)
        if b.dtype != a.dtype and a.dtype == DT:
            b = b.to(DT)
        elif b.dtype != a.dtype and b.dtype == DT:
            a = a.to(DT)
        return Var(fn(a, b))

    def __add__(self, o):
        return self._bin(o, torch.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __rsub__(self, o):
        return Var(_raw(o)).__sub__(self)

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, torch.div)

    def __rtruediv__(self, o):
        return Var(_raw(o)).__truediv__(self)

    def __pow__(self, o):
        return Var(self.v ** o)

    def __neg__(self):
        return Var(-self.v)

    def __matmul__(self, o):
        return self.dot(o)

    def dot(self, o):
        a, b = self.v, _raw(o)
        if a.ndim == 0 or b.ndim == 0:
            return Var(a * b)
        return Var(a @ b)

    # --- shape ops ---
    def __getitem__(self, idx):
        ridx = tuple(_raw(i) if isinstance(i, Var) else i for i in idx) if isinstance(idx, tuple) else (
            _raw(idx) if isinstance(idx, Var) else idx
        )
        if isinstance(ridx, slice) and ridx.step is not None and ridx.step < 0:
            assert ridx.start is None and ridx.stop is None and ridx.step == -1
            return Var(torch.flip(self.v, dims=(0,)))
        return Var(self.v[ridx], sub=(self, ridx))

    def astype(self, dtype):
        return Var(self.v.to(DT) if "float" in str(dtype) else self.v)

    def sum(self, axis=None):
        v = self.v.to(DT) if self.v.dtype == torch.bool else self.v
        return Var(v.sum() if axis is None else v.sum(dim=axis))

    def ravel(self):
        return Var(self.v.reshape(-1))

    def squeeze(self):
        return Var(self.v.squeeze())

    def copy(self):
        return Var(self.v.clone())


def _wrap1(fn):
    return lambda x, *a, **k: Var(fn(_raw(x), *a, **k))


def _scan(fn, sequences=None, outputs_info=None, non_sequences=None, go_backwards=False, name=None,
          mode=None, strict=False, **kw):
    sequences = [s if isinstance(s, Var) else Var(_raw(s)) for s in (sequences or [])]
    non_sequences = list(non_sequences or [])
    n = min(s.shape[0] for s in sequences)
    # PyTensor's scan() reverses every sequence FIRST when go_backwards is set (nw_seq = nw_seq[::-1]) and only then
    # levels them to the shortest length (scan_seqs = [seq[:actual_n_steps] ...], pytensor/scan/basic.py -- restated
    # from memory, PyTensor is not installable here).  For sequences of equal length this is plain reverse iteration;
    # for the smoother's unequal lengths (filtered[:-1] vs a time-varying T) it pairs filtered[t] with T[t+1].
    if go_backwards:
        sequences = [Var(torch.flip(s.v, dims=(0,))[:n]) for s in sequences]
    order = range(n)
    rec_idx = [i for i, o in enumerate(outputs_info) if o is not None]
    carry = [outputs_info[i] for i in rec_idx]
    collected = [[] for _ in outputs_info]
    for t in order:
        args = [Var(s.v[t]) for s in sequences] + list(carry) + non_sequences
        outs = fn(*args)
        outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
        for lst, o in zip(collected, outs):
            lst.append(_raw(o))
        carry = [outs[i] if isinstance(outs[i], Var) else Var(_raw(outs[i])) for i in rec_idx]
    results = [Var(torch.stack([x.to(DT) if x.dtype != DT else x for x in lst])) for lst in collected]
    return results, {}


class _PosSolve(torch.autograd.Function):
    """scipy.linalg.solve(A, b, assume_a='pos', lower=False) forward + PyTensor SolveBase.L_op backward."""

    @staticmethod
    def forward(ctx, A, b):
        Au = torch.triu(A) + torch.triu(A, 1).T
        c = torch.linalg.solve(Au, b)
        ctx.save_for_backward(Au, c)
        return c

    @staticmethod
    def backward(ctx, c_bar):
        Au, c = ctx.saved_tensors
        b_bar = torch.linalg.solve(Au.T, c_bar)
        A_bar = -torch.outer(b_bar, c) if c.ndim == 1 else -(b_bar @ c.T)
        return A_bar, b_bar


def _solve(A, b, assume_a="gen", check_finite=False, lower=False, **kw):
    A, b = _raw(A), _raw(b)
    if assume_a == "pos":
        return Var(_PosSolve.apply(A, b))
    return Var(torch.linalg.solve(A, b))


def _solve_triangular(A, b, lower=False, trans=0, unit_diagonal=False, check_finite=False, **kw):
    A, b = _raw(A), _raw(b)
    vec = b.ndim == 1
    bb = b[:, None] if vec else b
    x = torch.linalg.solve_triangular(A, bb, upper=not lower)
    return Var(x[:, 0] if vec else x)


def _cholesky(A, lower=True, **kw):
    A = _raw(A)
    # ?potrf(lower) reads the lower triangle only; differentiating through the explicit fill gives PyTensor's
    # Cholesky.L_op convention (tril(s + s.T) - diag(s)).  cholesky_ex never raises: PyTensor's ifelse is
    # lazy, ours is eager.
    A = torch.tril(A) + torch.tril(A, -1).T
    L, _info = torch.linalg.cholesky_ex(A)
    return Var(L if lower else L.T)


def _qr(A, mode="reduced"):
    Q, R = torch.linalg.qr(_raw(A), mode="reduced")
    if mode == "r":
        return Var(R)
    return Var(Q), Var(R)


def _pinv(A, **kw):
    return Var(torch.from_numpy(np.linalg.pinv(_raw(A).detach().numpy())))


def _set_subtensor(x, val):
    base, idx = x._sub
    out = base.v.clone()
    out[idx] = _raw(val).to(out.dtype) if not isinstance(val, (int, float)) else val
    return Var(out)


def _switch(cond, a, b):
    c = _raw(cond)
    if c.dtype != torch.bool:
        c = c != 0
    a, b = _raw(a), _raw(b)
    if a.dtype != b.dtype:
        a, b = a.to(DT), b.to(DT)
    return Var(torch.where(c, a, b))


def _ifelse(cond, a, b):
    c = bool(_raw(cond).item())
    return a if c else b


def _constant_fold(xs, raise_not_constant=True):
    return [tuple(int(s) for s in x) if isinstance(x, tuple) else x for x in xs]


class _Assert:
    def __init__(self, msg=""):
        self.msg = msg

    def __call__(self, value, *conds):
        for c in conds:
            assert bool(_raw(c).all()), self.msg
        return value


def _diag(x):
    return Var(torch.diag(_raw(x)))


def _tensor(name=None, dtype=None, shape=None):
    raise NotImplementedError("symbolic inputs are not part of the eager shim")


def install():
    """Register the stand-in modules in sys.modules (idempotent)."""
    if "pytensor" in sys.modules and getattr(sys.modules["pytensor"], "__ptshim__", False):
        return sys.modules["pytensor"]

    pytensor = types.ModuleType("pytensor")
    pytensor.__ptshim__ = True
    pytensor.config = types.SimpleNamespace(floatX="float64")
    pytensor.scan = _scan
    pytensor.ifelse = _ifelse

    pt = types.ModuleType("pytensor.tensor")
    pt.TensorVariable = Var
    pt.log = _wrap1(torch.log)
    pt.abs = _wrap1(torch.abs)
    pt.isnan = _wrap1(torch.isnan)
    pt.bitwise_not = _wrap1(torch.logical_not)
    pt.all = lambda x: Var(_raw(x).all())
    pt.constant = lambda x, dtype=None, **k: Var(torch.tensor(x, dtype=DT))
    pt.eq = lambda a, b: Var(_raw(a) == _raw(b))
    pt.neq = lambda a, b: Var(_raw(a) != _raw(b))
    pt.or_ = lambda a, b: Var(_raw(a) | _raw(b))
    pt.diag = _diag
    pt.eye = lambda n, m=None, k=0, dtype=None: Var(torch.eye(int(n), dtype=DT))
    pt.zeros = lambda shape, dtype=None: Var(torch.zeros(tuple(int(s) for s in shape) if shape != () else (), dtype=DT))
    pt.identity_like = lambda x: Var(torch.eye(_raw(x).shape[-1], dtype=DT))
    pt.set_subtensor = _set_subtensor
    pt.switch = _switch
    pt.horizontal_stack = lambda *xs: Var(torch.cat([_raw(x) for x in xs], dim=1))
    pt.vertical_stack = lambda *xs: Var(torch.cat([_raw(x) for x in xs], dim=0))
    pt.concatenate = lambda xs, axis=0: Var(torch.cat([_raw(x) for x in xs], dim=axis))
    pt.expand_dims = lambda x, axis: Var(_raw(x).unsqueeze(axis[0] if isinstance(axis, tuple) else axis))
    pt.specify_shape = lambda x, shape: x if isinstance(x, Var) else Var(_raw(x))
    pt.outer = lambda a, b: Var(torch.outer(_raw(a), _raw(b)))
    pt.atleast_1d = lambda x: Var(torch.atleast_1d(_raw(x)))
    pt.atleast_2d = lambda x: Var(torch.atleast_2d(_raw(x)))
    pt.einsum = lambda spec, *xs: Var(torch.einsum(spec, *[_raw(x) for x in xs]))
    pt.tensor = _tensor
    pt.linalg = types.SimpleNamespace(solve=_solve, det=_wrap1(torch.linalg.det), cholesky=_cholesky, qr=_qr,
                                      pinv=_pinv)

    slinalg = types.ModuleType("pytensor.tensor.slinalg")
    slinalg.solve_triangular = _solve_triangular
    nlinalg = types.ModuleType("pytensor.tensor.nlinalg")

    def matrix_dot(*args):
        out = args[0]
        for a in args[1:]:
            out = out.dot(a) if isinstance(out, Var) else Var(_raw(out)).dot(a)
        return out

    nlinalg.matrix_dot = matrix_dot

    compile_ = types.ModuleType("pytensor.compile")
    mode = types.ModuleType("pytensor.compile.mode")
    mode.get_mode = lambda m=None: m
    compile_.mode = mode
    compile_.get_mode = mode.get_mode
    graph = types.ModuleType("pytensor.graph")
    basic = types.ModuleType("pytensor.graph.basic")
    basic.Variable = Var
    graph.basic = basic
    raise_op = types.ModuleType("pytensor.raise_op")
    raise_op.Assert = _Assert

    pymc = types.ModuleType("pymc")
    pymc.__ptshim__ = True
    pytensorf = types.ModuleType("pymc.pytensorf")
    pytensorf.constant_fold = _constant_fold
    pymc.pytensorf = pytensorf

    pytensor.tensor = pt
    pt.slinalg, pt.nlinalg = slinalg, nlinalg
    pytensor.compile, pytensor.graph, pytensor.raise_op = compile_, graph, raise_op
    mods = {
        "pytensor": pytensor, "pytensor.tensor": pt, "pytensor.tensor.slinalg": slinalg,
        "pytensor.tensor.nlinalg": nlinalg, "pytensor.compile": compile_, "pytensor.compile.mode": mode,
        "pytensor.graph": graph, "pytensor.graph.basic": basic, "pytensor.raise_op": raise_op,
        "pymc": pymc, "pymc.pytensorf": pytensorf,
    }
    sys.modules.update(mods)
    return pytensor


def load_reference_filters(reference_root="/root/reference"):
    """Execute the reference's four hot-path source files in place and return
    (kalman_filter module, kalman_smoother module)."""
    install()
    base = f"{reference_root}/pymc_extras/statespace"
    # empty package stubs so that `import pymc_extras` does not execute the (PyMC-dependent) package __init__s
    for pkg in ("pymc_extras", "pymc_extras.statespace", "pymc_extras.statespace.filters",
                "pymc_extras.statespace.utils"):
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = []  # mark as package
            sys.modules[pkg] = mod

    def _load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("pymc_extras.statespace.utils.constants", f"{base}/utils/constants.py")
    _load("pymc_extras.statespace.filters.utilities", f"{base}/filters/utilities.py")
    kf = _load("pymc_extras.statespace.filters.kalman_filter", f"{base}/filters/kalman_filter.py")
    ks = _load("pymc_extras.statespace.filters.kalman_smoother", f"{base}/filters/kalman_smoother.py")
    return kf, ks


def to_var(x, requires_grad=False):
    t = torch.tensor(np.asarray(x, dtype=np.float64), dtype=DT, requires_grad=requires_grad)
    return Var(t)
