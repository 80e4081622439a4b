"""Low-level building blocks, same names and meaning as the reference's `algorithms` module
(src/algorithms/mod.rs, src/algorithms/lanczos.rs, src/algorithms/lanczos_two_pass.rs)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import STEP_CB, c_dp
from .error import LanczosError
from .operators import LinOp


@dataclass
class LanczosDecomposition:  # src/algorithms/mod.rs:94-108
    alphas: np.ndarray
    betas: np.ndarray
    steps_taken: int
    b_norm: float


@dataclass
class LanczosOutput:  # src/algorithms/mod.rs:115-122
    v_k: np.ndarray  # n x steps_taken, column-major
    decomposition: LanczosDecomposition


@dataclass
class LanczosPassTwoOutput:  # src/algorithms/mod.rs:130-135
    x_k: np.ndarray
    v_k: np.ndarray


@dataclass
class TridiagonalSystemView:  # src/algorithms/mod.rs:57-67
    alphas: np.ndarray
    betas: np.ndarray
    steps_taken: int


def _dp(a):
    return a.ctypes.data_as(c_dp)


def lanczos_pass_one(operator: LinOp, b, k: int) -> LanczosDecomposition:
    """src/algorithms/lanczos_two_pass.rs:65-110"""
    bp, keep, _ = operator._vec(b)
    al = np.zeros(max(k, 1))
    be = np.zeros(max(k, 1))
    steps, bn = C.c_size_t(), C.c_double()
    _lib.check(_lib.load().tpl_pass_one(operator._h, bp, k, _dp(al), _dp(be), C.byref(steps), C.byref(bn)))
    s = steps.value
    return LanczosDecomposition(al[:s].copy(), be[:max(s - 1, 0)].copy(), s, bn.value)


def _pass_two(operator: LinOp, b, dec: LanczosDecomposition, y_k, with_basis: bool):
    bp, keep, is_torch = operator._vec(b)
    n = operator.nrows()
    y = np.ascontiguousarray(np.asarray(y_k, dtype=np.float64).reshape(-1))
    al = np.ascontiguousarray(dec.alphas, dtype=np.float64)
    be = np.ascontiguousarray(dec.betas, dtype=np.float64)
    # the C entry point reads steps_taken alphas and steps_taken - 1 betas through bare pointers
    if len(al) < dec.steps_taken:
        raise LanczosError(4, f"Parameter mismatch: `alphas` expects size {dec.steps_taken}, but got {len(al)}.")
    if len(be) < max(dec.steps_taken - 1, 0):
        raise LanczosError(4, f"Parameter mismatch: `betas` expects size {max(dec.steps_taken - 1, 0)}, but got {len(be)}.")
    al = al if len(al) else np.zeros(1)
    be = be if len(be) else np.zeros(1)
    y_arg = y if len(y) else np.zeros(1)
    V = np.zeros((dec.steps_taken, n)) if with_basis else None  # row j of this array == column j of V'
    if is_torch and keep.is_cuda:
        x = keep.new_empty(n)
        xptr = C.c_void_p(x.data_ptr())
    else:
        x = np.empty(n)
        xptr = C.c_void_p(x.ctypes.data)
    _lib.check(_lib.load().tpl_pass_two(operator._h, bp, _dp(al), _dp(be), dec.steps_taken, dec.b_norm, _dp(y_arg),
                                        len(y), xptr, C.c_void_p(V.ctypes.data) if with_basis and V.size else None,
                                        n))
    return x, (V.T if with_basis else None)


def lanczos_pass_two(operator: LinOp, b, decomposition: LanczosDecomposition, y_k):
    """src/algorithms/lanczos_two_pass.rs:128-140"""
    return _pass_two(operator, b, decomposition, y_k, False)[0]


def lanczos_pass_two_with_basis(operator: LinOp, b, decomposition: LanczosDecomposition, y_k) -> LanczosPassTwoOutput:
    """src/algorithms/lanczos_two_pass.rs:149-166"""
    x, v = _pass_two(operator, b, decomposition, y_k, True)
    return LanczosPassTwoOutput(x, v)


def lanczos_standard(operator: LinOp, b, k: int, callback=None) -> LanczosOutput:
    """src/algorithms/lanczos.rs:55-156.  `callback(k, v_k, t_k_view) -> bool` follows LanczosCallback
    (mod.rs:82-86); v_k is handed over as a lazily-fetched object with `.to_host()` (it lives in HBM)."""
    bp, keep, _ = operator._vec(b)
    n = operator.nrows()
    V = np.zeros((max(k, 1), n))
    al = np.zeros(max(k, 1))
    be = np.zeros(max(k, 1))
    steps, bn = C.c_size_t(), C.c_double()

    def _cb(s, vdev, ld, ap, bp_, _u):
        betas = np.ctypeslib.as_array(bp_, shape=(s - 1,)).copy() if s > 1 and bp_ else np.zeros(0)
        view = TridiagonalSystemView(np.ctypeslib.as_array(ap, shape=(s,)).copy(), betas, s)
        return 1 if callback(s, DeviceBasisView(vdev, ld, n, s), view) else 0

    cb = STEP_CB(_cb) if callback is not None else C.cast(None, STEP_CB)
    _lib.check(_lib.load().tpl_standard(operator._h, bp, k, C.c_void_p(V.ctypes.data), n, _dp(al), _dp(be),
                                        C.byref(steps), C.byref(bn), cb, None))
    s = steps.value
    dec = LanczosDecomposition(al[:s].copy(), be[:max(s - 1, 0)].copy(), s, bn.value)
    return LanczosOutput(V[:s].T, dec)  # trimmed to steps_taken columns (lanczos.rs:135-145)


class DeviceBasisView:
    """n x steps column-major basis in HBM as passed to a LanczosCallback."""

    def __init__(self, ptr, ld, n, steps):
        self.ptr, self.ld, self.n, self.steps = ptr, ld, n, steps

    def to_host(self) -> np.ndarray:
        import torch

        out = torch.empty((self.steps, self.ld), dtype=torch.float64)
        nbytes = self.steps * self.ld * 8
        rc = torch.cuda.cudart().cudaMemcpy(out.data_ptr(), self.ptr, nbytes, 2)  # cudaMemcpyDeviceToHost
        if int(rc) != 0:
            raise RuntimeError(f"cudaMemcpy failed: {rc}")
        return out.numpy()[:, : self.n].T
