"""Device-resident operators: the stand-in for `&impl faer::matrix_free::LinOp<f64>`
(src/solvers.rs:56, src/algorithms/mod.rs:167).  Thin RAII wrappers over `tpl_op*` handles."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, c_u32p, c_u64p


def _vec_ptr(v):
    """(pointer, keepalive, is_torch) for a host numpy vector or a torch tensor (host or CUDA)."""
    if hasattr(v, "data_ptr"):  # torch tensor; CUDA tensors are consumed in place (no copy)
        t = v.contiguous()
        if str(t.dtype) == "torch.complex128":  # n complex numbers = 2 n doubles, interleaved (Hermitian operators)
            import torch

            t = torch.view_as_real(t).reshape(-1)
        if str(t.dtype) != "torch.float64":
            raise TypeError("vectors must be float64 (complex128 for a Hermitian operator)")
        return C.c_void_p(t.data_ptr()), t, True
    a = np.asarray(v)
    if np.iscomplexobj(a):
        a = np.ascontiguousarray(a.astype(np.complex128, copy=False).reshape(-1)).view(np.float64)
    else:
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    return C.c_void_p(a.ctypes.data), a, False


class LinOp:
    """Owns a `tpl_op*`.  `nrows()`/`ncols()`/`apply()` mirror faer's LinOp trait."""

    FORMAT = {1: "csr", 2: "incidence", 3: "dense"}

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        self._stream = None  # None = the handle's own stream; else the cudaStream_t it was given

    def _vec(self, v):
        """`_vec_ptr` plus the two checks a bare pointer cannot carry: the vector has nrows() entries (else the reference's
        DimensionMismatch, src/error.rs:29-35) and, for a CUDA tensor, the handle works on torch's CURRENT stream so that
        it is ordered after the kernels that produced the tensor and before the ones that will consume the result."""
        if self.is_complex and not hasattr(v, "data_ptr") and not np.iscomplexobj(v) and np.size(v) * 2 == self.nrows():
            v = np.asarray(v, dtype=np.complex128)  # a real vector for a Hermitian operator
        ptr, keep, is_torch = _vec_ptr(v)
        _lib.check(_lib.load().tpl_op_check_len(self._h, int(keep.numel()) if is_torch else int(keep.shape[0])))
        if is_torch and keep.is_cuda:
            import torch

            cur = int(torch.cuda.current_stream(keep.device).cuda_stream)
            if self._stream != cur:
                self.set_stream(cur)
        return ptr, keep, is_torch

    @property
    def is_complex(self) -> bool:
        """Hermitian operator: vectors are complex128 arrays (stored interleaved, nrows() = 2 x their length)."""
        if getattr(self, "_cplx", None) is None:
            self._cplx = bool(_lib.load().tpl_op_is_complex(self._h))
        return self._cplx

    def _out(self, x):
        """Result vectors of a Hermitian operator go back as complex128 views of the interleaved storage."""
        if not self.is_complex:
            return x
        if hasattr(x, "data_ptr"):
            import torch

            return torch.view_as_complex(x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2))
        return x.view(np.complex128)

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_csc(cls, n, colptr, rowidx, val, device: int = -1) -> "LinOp":
        colptr = np.ascontiguousarray(colptr, dtype=np.uint64)
        rowidx = np.ascontiguousarray(rowidx, dtype=np.uint64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(_lib.load().tpl_op_from_csc(n, colptr.ctypes.data_as(c_u64p), rowidx.ctypes.data_as(c_u64p),
                                               val.ctypes.data_as(c_dp), device, C.byref(h)))
        return cls(h)

    @classmethod
    def from_scipy(cls, a, device: int = -1) -> "LinOp":
        import scipy.sparse as sp

        a = sp.csc_matrix(a)
        a.sort_indices()
        return cls.from_csc(a.shape[0], a.indptr, a.indices, a.data, device)

    @classmethod
    def from_dense(cls, a, device: int = -1) -> "LinOp":
        """Dense symmetric operator (faer `Mat<f64>` as LinOp, src/bin/dense_tradeoff.rs:154-162): streamed by the dense
        kernels, 8 n^2 bytes per product."""
        a = np.asfortranarray(np.asarray(a, dtype=np.float64))
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("a square matrix is required")
        h = C.c_void_p()
        _lib.check(_lib.load().tpl_op_from_dense(a.shape[0], a.ctypes.data_as(c_dp), a.shape[0], device, C.byref(h)))
        return cls(h)

    @classmethod
    def from_dense_hermitian(cls, a, device: int = -1) -> "LinOp":
        """Dense complex Hermitian operator (`T: ComplexField`, src/algorithms/mod.rs:167): vectors in and out are complex128,
        alpha / beta stay real.  16 n^2 bytes per product."""
        a = np.asfortranarray(np.asarray(a, dtype=np.complex128))
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("a square matrix is required")
        h = C.c_void_p()
        _lib.check(_lib.load().tpl_op_from_dense_hermitian(a.shape[0], C.cast(a.ctypes.data, c_dp), a.shape[0], device,
                                                           C.byref(h)))
        return cls(h)

    @classmethod
    def from_diagonal(cls, diag, device: int = -1) -> "LinOp":
        """diag(d): the synthetic spectra of src/bin/stability.rs / orthogonality.rs."""
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(_lib.load().tpl_op_from_diagonal(len(diag), diag.ctypes.data_as(c_dp), device, C.byref(h)))
        return cls(h)

    @classmethod
    def from_kkt(cls, m, p, tail, head, d, device: int = -1) -> "LinOp":
        tail = np.ascontiguousarray(tail, dtype=np.uint32)
        head = np.ascontiguousarray(head, dtype=np.uint32)
        d = np.ascontiguousarray(d, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(_lib.load().tpl_op_from_kkt(m, p, tail.ctypes.data_as(c_u32p), head.ctypes.data_as(c_u32p),
                                               d.ctypes.data_as(c_dp), len(d), device, C.byref(h)))
        return cls(h)

    # -- LinOp ------------------------------------------------------------------------------------
    def nrows(self) -> int:
        return _lib.load().tpl_op_nrows(self._h)

    ncols = nrows

    @property
    def format(self) -> str:
        return self.FORMAT.get(_lib.load().tpl_op_format(self._h), "?")

    def apply(self, x):
        xp, keep, is_torch = self._vec(x)
        if is_torch and keep.is_cuda:
            y = keep.new_empty(keep.shape)
            _lib.check(_lib.load().tpl_op_apply(self._h, xp, C.c_void_p(y.data_ptr())))
            return self._out(y)
        y = np.empty(self.nrows())
        _lib.check(_lib.load().tpl_op_apply(self._h, xp, C.c_void_p(y.ctypes.data)))
        return self._out(y)

    # -- engine knobs / introspection -----------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        _lib.check(_lib.load().tpl_op_set_stream(self._h, C.c_void_p(cuda_stream)))
        self._stream = int(cuda_stream)

    def set_mode(self, mode: int):
        _lib.check(_lib.load().tpl_op_set_mode(self._h, mode))

    def kernel_shape(self) -> str:
        """kernel family a whole-pass solve runs: cells / chunks / tiled / gather / csr / sharded"""
        return _lib.load().tpl_op_kernel_shape(self._h).decode()

    def layout_check(self, tail, head, d) -> dict:
        """Downloads the handle's tables (built on the device from 2^20 arcs on) and checks them on the host
        (tpl_op_layout_check): blocked layout consistent, node lists equal to a host construction."""
        tail = np.ascontiguousarray(tail, dtype=np.uint32)
        head = np.ascontiguousarray(head, dtype=np.uint32)
        d = np.ascontiguousarray(d, dtype=np.float64)
        st = (C.c_uint64 * 8)()
        _lib.check(_lib.load().tpl_op_layout_check(self._h, tail.ctypes.data_as(c_u32p), head.ctypes.data_as(c_u32p),
                                                   d.ctypes.data_as(c_dp), len(d), st))
        keys = ("device_built", "blocked", "check", "hash", "list_words", "tile_arcs", "node_list_mismatches", "node_list_words")
        return dict(zip(keys, (int(x) for x in st)))

    def trace_enable(self, max_steps: int):
        _lib.check(_lib.load().tpl_op_trace_enable(self._h, max_steps))

    def trace_read(self) -> np.ndarray:
        """uint64[ctas, steps, marks] phase timestamps of the last resident-kernel launches."""
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.check(_lib.load().tpl_op_trace_read(self._h, None, 0, C.byref(a), C.byref(b), C.byref(c)))
        out = np.zeros((a.value, b.value, c.value), dtype=np.uint64)
        if out.size:
            _lib.check(_lib.load().tpl_op_trace_read(self._h, out.ctypes.data_as(_lib.c_u64p), out.size, C.byref(a),
                                                     C.byref(b), C.byref(c)))
        return out

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _lib.check(_lib.load().tpl_op_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"pass_one_ms": a.value, "pass_two_ms": b.value, "gemv_ms": c.value}

    def fabric_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's exchange block (sharded operators)."""
        buf = (C.c_uint8 * 64)()
        _lib.check(_lib.load().tpl_op_fabric_export(self._h, buf))
        return bytes(buf)

    def fabric_import(self, handles):
        """Maps the exchange blocks of all ranks (their handles in rank order): the passes then run fused."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _lib.check(_lib.load().tpl_op_fabric_import(self._h, buf, len(handles)))

    def shard_info(self):
        r, w, a, p = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        _lib.check(_lib.load().tpl_op_shard_info(self._h, C.byref(r), C.byref(w), C.byref(a), C.byref(p)))
        return {"rank": r.value, "world": w.value, "local_arcs": a.value, "nodes": p.value}

    def kernel_launches(self) -> int:
        return _lib.load().tpl_op_kernel_launches(self._h)

    def matrix_bytes(self) -> int:
        return _lib.load().tpl_op_matrix_bytes(self._h)

    def device_bytes(self) -> int:
        return _lib.load().tpl_op_device_bytes(self._h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tpl_op_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
