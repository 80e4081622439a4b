"""`.dmx` / `.qfc` KKT loader, same names as the reference's utils::data_loader (src/utils/data_loader.rs)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import c_dp, c_u32p, c_u64p
from .operators import LinOp


@dataclass
class KKTSystem:  # src/utils/data_loader.rs:51-58
    a: LinOp           # device-resident operator standing in for SparseColMat<usize, f64>
    num_nodes: int
    num_arcs: int
    host: "HostKKT"    # the host-side matrix (CSC + incidence views), kept for tests / export


class HostKKT:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.load().tpl_kkt_free(self._h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    @property
    def num_nodes(self):
        return _lib.load().tpl_kkt_num_nodes(self._h)

    @property
    def num_arcs(self):
        return _lib.load().tpl_kkt_num_arcs(self._h)

    @property
    def num_costs(self):
        return _lib.load().tpl_kkt_num_costs(self._h)

    def csc(self):
        """(n, colptr u64[n+1], rowidx u64[nnz], val f64[nnz]) copies of KKTSystem.a"""
        n, nnz = C.c_size_t(), C.c_size_t()
        cp, ri, va = c_u64p(), c_u64p(), c_dp()
        _lib.check(_lib.load().tpl_kkt_csc(self._h, C.byref(n), C.byref(nnz), C.byref(cp), C.byref(ri), C.byref(va)))
        colptr = np.ctypeslib.as_array(cp, shape=(n.value + 1,)).copy()
        if nnz.value == 0:
            return n.value, colptr, np.zeros(0, np.uint64), np.zeros(0)
        rowidx = np.ctypeslib.as_array(ri, shape=(nnz.value,)).copy()
        val = np.ctypeslib.as_array(va, shape=(nnz.value,)).copy()
        return n.value, colptr, rowidx, val

    def incidence(self):
        """(tail u32[m], head u32[m], d f64[m] zero padded, d_len, regular)"""
        t, h, d = c_u32p(), c_u32p(), c_dp()
        dl, reg = C.c_size_t(), C.c_int()
        _lib.check(_lib.load().tpl_kkt_incidence(self._h, C.byref(t), C.byref(h), C.byref(d), C.byref(dl), C.byref(reg)))
        m = self.num_arcs
        if m == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0), dl.value, bool(reg.value)
        sh = (m,)
        return (np.ctypeslib.as_array(t, shape=sh)[:m].copy(), np.ctypeslib.as_array(h, shape=sh)[:m].copy(),
                np.ctypeslib.as_array(d, shape=sh)[:m].copy(), dl.value, bool(reg.value))


    def save_binary(self, path):
        """Writes the instance as a TPLKKT1 container (include/tplanczos.h)."""
        _lib.check(_lib.load().tpl_kkt_save_binary(self._h, os.fsencode(path)))


def write_kkt_binary(path, num_nodes, tail, head, costs=()):
    """Container straight from an arc list (0-based tail/head); `costs` may be shorter than the arc list (short D)."""
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    if tail.shape != head.shape:
        raise ValueError("tail and head must have the same length")
    _lib.check(_lib.load().tpl_write_kkt_binary(os.fsencode(path), num_nodes, len(tail), tail.ctypes.data_as(c_u32p),
                                                head.ctypes.data_as(c_u32p), costs.ctypes.data_as(c_dp), len(costs)))


def load_kkt_host_binary(path) -> HostKKT:
    h = C.c_void_p()
    _lib.check(_lib.load().tpl_load_kkt_binary(os.fsencode(path), C.byref(h)))
    return HostKKT(h)


def load_kkt_system_binary(path, fmt: str = "auto", device: int = -1) -> KKTSystem:
    """`load_kkt_system` for a TPLKKT1 container."""
    host = load_kkt_host_binary(path)
    h = C.c_void_p()
    code = {"auto": 0, "csr": 1, "incidence": 2}[fmt]
    _lib.check(_lib.load().tpl_op_from_kkt_system(host._h, code, device, C.byref(h)))
    return KKTSystem(LinOp(h), host.num_nodes, host.num_arcs, host)


def load_kkt_host(dmx_path, qfc_path) -> HostKKT:
    h = C.c_void_p()
    _lib.check(_lib.load().tpl_load_kkt(os.fsencode(dmx_path), os.fsencode(qfc_path), C.byref(h)))
    return HostKKT(h)


def load_kkt_system(dmx_path, qfc_path, fmt: str = "auto", device: int = -1) -> KKTSystem:
    """src/utils/data_loader.rs:211-259.  fmt: "auto" | "csr" | "incidence" selects the kernel-side format."""
    host = load_kkt_host(dmx_path, qfc_path)
    h = C.c_void_p()
    code = {"auto": 0, "csr": 1, "incidence": 2}[fmt]
    _lib.check(_lib.load().tpl_op_from_kkt_system(host._h, code, device, C.byref(h)))
    return KKTSystem(LinOp(h), host.num_nodes, host.num_arcs, host)
