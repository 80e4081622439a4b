"""ctypes front-end of the CPU oracle (oracle/lanczos_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(two_pass_lanczos_b200/) never imports this module.

Function names and argument meaning follow the reference's public API
(src/solvers.rs:46-175, src/algorithms/lanczos.rs:55, src/algorithms/lanczos_two_pass.rs:65-166,
src/utils/data_loader.rs:211) so that tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblanczos_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (see oracle/Makefile)."""
    src = os.path.join(_HERE, "lanczos_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class OracleError(Exception):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        self.message = message


_FTK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double), C.c_size_t,
                   C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.c_void_p)
_STEPCB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, sp, up = C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)
        vp = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_csc_new.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, up, up, dp, C.POINTER(vp)]
        L.orc_csc_free.argtypes = [vp]
        L.orc_csc_nrows.argtypes = [vp]
        L.orc_csc_nrows.restype = C.c_size_t
        L.orc_csc_nnz.argtypes = [vp]
        L.orc_csc_nnz.restype = C.c_size_t
        L.orc_csc_export.argtypes = [vp, up, up, dp]
        L.orc_load_kkt.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), sp, sp]
        L.orc_matvec.argtypes = [vp, dp, dp]
        L.orc_pass_one.argtypes = [vp, dp, C.c_size_t, dp, dp, sp, dp]
        L.orc_pass_one_ld.argtypes = [vp, dp, C.c_size_t, dp, dp, sp, dp]
        L.orc_pass_two.argtypes = [vp, dp, dp, dp, C.c_size_t, C.c_double, dp, C.c_size_t, dp, dp]
        L.orc_standard.argtypes = [vp, dp, C.c_size_t, dp, dp, dp, sp, dp, _STEPCB, vp]
        L.orc_lanczos.argtypes = [vp, dp, C.c_size_t, _FTK, vp, dp]
        L.orc_lanczos_two_pass.argtypes = [vp, dp, C.c_size_t, _FTK, vp, dp]
        L.orc_lanczos_two_pass_ld.argtypes = [vp, dp, C.c_size_t, _FTK, vp, dp]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


class SparseColMat:
    """faer::sparse::SparseColMat<usize,f64> stand-in (CSC, 8-byte indices)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def try_new_from_triplets(cls, nrows, ncols, rows, cols, vals) -> "SparseColMat":
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        vals = _f64(vals)
        h = C.c_void_p()
        up = C.POINTER(C.c_uint64)
        _check(lib().orc_csc_new(nrows, ncols, len(vals), rows.ctypes.data_as(up), cols.ctypes.data_as(up),
                                 _dp(vals), C.byref(h)))
        return cls(h)

    @classmethod
    def from_dense(cls, a: np.ndarray) -> "SparseColMat":
        a = np.asarray(a, dtype=np.float64)
        r, c = np.nonzero(a)
        return cls.try_new_from_triplets(a.shape[0], a.shape[1], r, c, a[r, c])

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_csc_free(self._h)
            self._h = None

    def nrows(self) -> int:
        return lib().orc_csc_nrows(self._h)

    def nnz(self) -> int:
        return lib().orc_csc_nnz(self._h)

    def csc(self):
        """(colptr u64[n+1], rowidx u64[nnz], val f64[nnz])"""
        n, nnz = self.nrows(), self.nnz()
        colptr = np.empty(n + 1, np.uint64)
        rowidx = np.empty(nnz, np.uint64)
        val = np.empty(nnz, np.float64)
        up = C.POINTER(C.c_uint64)
        lib().orc_csc_export(self._h, colptr.ctypes.data_as(up), rowidx.ctypes.data_as(up), _dp(val))
        return colptr, rowidx, val

    def apply(self, x) -> np.ndarray:
        x = _f64(x)
        y = np.empty_like(x)
        lib().orc_matvec(self._h, _dp(x), _dp(y))
        return y


@dataclass
class KKTSystem:  # src/utils/data_loader.rs:51-58
    a: SparseColMat
    num_nodes: int
    num_arcs: int


def load_kkt_system(dmx_path, qfc_path) -> KKTSystem:
    h = C.c_void_p()
    nodes, arcs = C.c_size_t(), C.c_size_t()
    _check(lib().orc_load_kkt(os.fsencode(dmx_path), os.fsencode(qfc_path), C.byref(h), C.byref(nodes),
                              C.byref(arcs)))
    return KKTSystem(SparseColMat(h), nodes.value, arcs.value)


@dataclass
class LanczosDecomposition:  # src/algorithms/mod.rs:94-108
    alphas: np.ndarray
    betas: np.ndarray
    steps_taken: int
    b_norm: float


def lanczos_pass_one(op: SparseColMat, b, k: int, extended: bool = False) -> LanczosDecomposition:
    b = _f64(b)
    al = np.zeros(max(k, 1))
    be = np.zeros(max(k, 1))
    steps, bn = C.c_size_t(), C.c_double()
    fn = lib().orc_pass_one_ld if extended else lib().orc_pass_one
    _check(fn(op._h, _dp(b), k, _dp(al), _dp(be), C.byref(steps), C.byref(bn)))
    s = steps.value
    return LanczosDecomposition(al[:s].copy(), be[:max(s - 1, 0)].copy(), s, bn.value)


def lanczos_pass_two(op, b, dec: LanczosDecomposition, y_k, with_basis: bool = False):
    b, y = _f64(b), _f64(y_k)
    n = op.nrows()
    x = np.empty(n)
    V = np.zeros((dec.steps_taken, n)) if with_basis else None  # row j = column j (col-major n x steps)
    al, be = _f64(dec.alphas), _f64(dec.betas)
    if len(be) == 0:
        be = np.zeros(1)
    if len(al) == 0:
        al = np.zeros(1)
    _check(lib().orc_pass_two(op._h, _dp(b), _dp(al), _dp(be), dec.steps_taken, dec.b_norm, _dp(y), len(y),
                              _dp(x), _dp(V) if with_basis else None))
    return (x, V.T) if with_basis else x


def lanczos_standard(op, b, k: int, callback=None):
    """returns (V n x steps, LanczosDecomposition)"""
    b = _f64(b)
    n = op.nrows()
    V = np.zeros((max(k, 1), n))
    al = np.zeros(max(k, 1))
    be = np.zeros(max(k, 1))
    steps, bn = C.c_size_t(), C.c_double()

    def _cb(s, vptr, ld, ap, bp, _u):
        v = np.ctypeslib.as_array(vptr, shape=(s, ld)).T
        a_ = np.ctypeslib.as_array(ap, shape=(s,))
        b_ = np.ctypeslib.as_array(bp, shape=(max(s - 1, 1),))[:s - 1]
        return 1 if callback(s, v, a_, b_) else 0

    cb = _STEPCB(_cb) if callback is not None else C.cast(None, _STEPCB)
    _check(lib().orc_standard(op._h, _dp(b), k, _dp(V), _dp(al), _dp(be), C.byref(steps), C.byref(bn), cb, None))
    s = steps.value
    dec = LanczosDecomposition(al[:s].copy(), be[:max(s - 1, 0)].copy(), s, bn.value)
    return V[:s].T.copy(order="F"), dec


def _wrap_ftk(f_tk_solver, err_box):
    def _f(ap, na, bp, nb, yp, ylenp, _u):
        try:
            a_ = np.ctypeslib.as_array(ap, shape=(na,)).copy() if na else np.zeros(0)
            b_ = np.ctypeslib.as_array(bp, shape=(nb,)).copy() if nb else np.zeros(0)
            y = np.asarray(f_tk_solver(a_, b_), dtype=np.float64).reshape(-1)
            cap = na + 64
            ylenp[0] = len(y)
            for i in range(min(len(y), cap)):
                yp[i] = y[i]
            return 0
        except Exception as e:  # noqa: BLE001 - closure errors become SolverError (solvers.rs:72,156)
            err_box.append(e)
            return 1

    return _FTK(_f)


def lanczos(op, b, k, f_tk_solver) -> np.ndarray:
    b = _f64(b)
    x = np.empty(op.nrows())
    box: list = []
    _check(lib().orc_lanczos(op._h, _dp(b), k, _wrap_ftk(f_tk_solver, box), None, _dp(x)))
    return x


def lanczos_two_pass(op, b, k, f_tk_solver, extended: bool = False) -> np.ndarray:
    b = _f64(b)
    x = np.empty(op.nrows())
    box: list = []
    fn = lib().orc_lanczos_two_pass_ld if extended else lib().orc_lanczos_two_pass
    _check(fn(op._h, _dp(b), k, _wrap_ftk(f_tk_solver, box), None, _dp(x)))
    return x
