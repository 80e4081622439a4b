"""High-level solvers, same names and meaning as the reference's `solvers` module (src/solvers.rs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FTK_FN
from .operators import LinOp

# Host f(T_k) e1 solvers implemented in the library (C++); usable wherever a closure is expected.
INV, EXP, SQUARE = "inv", "exp", "square"


def _native_ftk(name: str):
    fn = getattr(_lib.load(), {"inv": "tpl_ftk_inv", "exp": "tpl_ftk_exp", "square": "tpl_ftk_square"}[name])
    return C.cast(fn, FTK_FN)


def _wrap_closure(f_tk_solver, errors: list):
    """F: FnMut(&[R], &[R]) -> Result<Mat<T>, anyhow::Error>  (solvers.rs:58)"""

    def _f(ap, na, bp, nb, yp, ylenp, _u):
        try:
            a_ = np.ctypeslib.as_array(ap, shape=(na,)).copy() if na else np.zeros(0)
            b_ = np.ctypeslib.as_array(bp, shape=(nb,)).copy() if nb else np.zeros(0)
            y = np.asarray(f_tk_solver(a_, b_), dtype=np.float64)
            if y.ndim == 2 and y.shape[1] != 1:  # y' must be steps x 1 (solvers.rs:75,158)
                ylenp[0] = y.shape[0] + 1 if y.shape[0] == na else y.shape[0]
                return 0
            y = y.reshape(-1)
            ylenp[0] = len(y)
            for i in range(min(len(y), na)):
                yp[i] = y[i]
            return 0
        except Exception as e:  # noqa: BLE001 - becomes SolverError(e.to_string()) (solvers.rs:72,156)
            errors.append(e)
            return 1

    return FTK_FN(_f)


def _solve(entry: str, operator: LinOp, b, k: int, f_tk_solver):
    bp, keep, is_torch = operator._vec(b)
    n = operator.nrows()
    if is_torch and keep.is_cuda:
        x = keep.new_empty(n)
        xptr = C.c_void_p(x.data_ptr())
    elif is_torch and keep.is_pinned():
        # pinned host tensor in -> pinned host tensor out (torch caches pinned blocks): the result comes back in one DMA
        # instead of the driver's staged copy into pageable memory
        import torch

        x = torch.empty(n, dtype=torch.float64, pin_memory=True)
        xptr = C.c_void_p(x.data_ptr())
    else:
        x = np.empty(n)
        xptr = C.c_void_p(x.ctypes.data)
    errors: list = []
    cb = _native_ftk(f_tk_solver) if isinstance(f_tk_solver, str) else _wrap_closure(f_tk_solver, errors)
    rc = getattr(_lib.load(), entry)(operator._h, bp, k, cb, None, xptr)
    if rc == 6 and errors:  # SolverError carries the closure's message
        from .error import LanczosError

        raise LanczosError(6, f"The user-provided f(T_k) solver failed: {errors[0]}")
    _lib.check(rc)
    return operator._out(x)


def lanczos(operator: LinOp, b, k: int, f_tk_solver):
    """One-pass f(A)b (src/solvers.rs:46-107): V_k kept in HBM, x = ||b|| V_k f(T_k) e1."""
    return _solve("tpl_lanczos", operator, b, k, f_tk_solver)


def lanczos_two_pass(operator: LinOp, b, k: int, f_tk_solver):
    """Two-pass f(A)b (src/solvers.rs:133-175): O(n) memory, basis regenerated in pass 2."""
    return _solve("tpl_lanczos_two_pass", operator, b, k, f_tk_solver)


def _sweep(entry: str, operator: LinOp, b, ks, f_tk_solver):
    bp, keep, is_torch = operator._vec(b)
    n = operator.nrows()
    ks_arr = (C.c_size_t * len(ks))(*[int(k) for k in ks])
    if is_torch and keep.is_cuda:
        X = keep.new_empty((len(ks), n))          # row q = x_q (column-major n x nk for the library)
        xptr = C.c_void_p(X.data_ptr())
    else:
        X = np.empty((len(ks), n))
        xptr = C.c_void_p(X.ctypes.data)
    errors: list = []
    cb = _native_ftk(f_tk_solver) if isinstance(f_tk_solver, str) else _wrap_closure(f_tk_solver, errors)
    rc = getattr(_lib.load(), entry)(operator._h, bp, ks_arr, len(ks), cb, None, xptr, n)
    if rc == 6 and errors:
        from .error import LanczosError

        raise LanczosError(6, f"The user-provided f(T_k) solver failed: {errors[0]}")
    _lib.check(rc)
    return X


def lanczos_sweep(operator: LinOp, b, ks, f_tk_solver):
    """One-pass k-sweep: x_q = f(A) b with ks[q] steps for every q from ONE basis generation to max(ks) and one streaming pass
    over the basis (the reference re-solves per k, src/bin/tradeoff.rs:262-290).  Returns an array whose row q is bit-identical
    to `lanczos(operator, b, ks[q], f_tk_solver)`."""
    return _sweep("tpl_lanczos_sweep", operator, b, ks, f_tk_solver)


def lanczos_two_pass_sweep(operator: LinOp, b, ks, f_tk_solver):
    """Two-pass k-sweep with O(n) memory: ONE pass 1 to max(ks), one pass 2 per k.  Row q is bit-identical to
    `lanczos_two_pass(operator, b, ks[q], f_tk_solver)`."""
    return _sweep("tpl_lanczos_two_pass_sweep", operator, b, ks, f_tk_solver)


def inv_residual_estimates(alphas, betas, b_norm: float = 1.0):
    """||b - A x_j||, j = 1..len(alphas), of the iterates x_j = ||b|| V_j T_j^{-1} e_1, from the coefficients of one pass 1
    (`tpl_ftk_inv_residuals`).  `betas[j-1]` = beta_j; entries whose beta is not given come back NaN (a
    reference-style decomposition holds steps - 1 betas, so its last estimate is NaN)."""
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    be = np.ascontiguousarray(betas, dtype=np.float64)
    res = np.empty(len(al))
    dp = _lib.c_dp
    _lib.check(_lib.load().tpl_ftk_inv_residuals((al if len(al) else np.zeros(1)).ctypes.data_as(dp), len(al),
                                                 (be if len(be) else np.zeros(1)).ctypes.data_as(dp), len(be), b_norm,
                                                 (res if len(res) else np.zeros(1)).ctypes.data_as(dp)))
    return res


def lanczos_two_pass_inv_adaptive(operator: LinOp, b, k_max: int, rtol: float):
    """A x = b with the number of steps chosen from the residual estimates of ONE pass 1: returns
    (x, k_used, residual_estimate).  k_used is the first j <= k_max whose estimate is <= rtol ||b|| (else the best j);
    pass 2 regenerates only k_used basis vectors."""
    bp, keep, is_torch = operator._vec(b)
    n = operator.nrows()
    if is_torch and keep.is_cuda:
        x = keep.new_empty(n)
        xptr = C.c_void_p(x.data_ptr())
    else:
        x = np.empty(n)
        xptr = C.c_void_p(x.ctypes.data)
    k_used, est = C.c_size_t(), C.c_double()
    _lib.check(_lib.load().tpl_lanczos_two_pass_inv_adaptive(operator._h, bp, k_max, rtol, xptr, C.byref(k_used), C.byref(est)))
    return x, k_used.value, est.value
