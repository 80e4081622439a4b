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
        rowidx = np.ctypeslib.as_array(ri, shape=(max(nnz.value, 1),))[: nnz.value].copy()
        val = np.ctypeslib.as_array(va, shape=(max(nnz.value, 1),))[: nnz.value].copy()
        return n.value, colptr, rowidx, val

    def incidence(self):
        """(tail u32[m], head u32[m], d f64[m] zero padded, d_len, regular)"""
        t, h, d = c_u32p(), c_u32p(), c_dp()
        dl, reg = C.c_size_t(), C.c_int()
        _lib.check(_lib.load().tpl_kkt_incidence(self._h, C.byref(t), C.byref(h), C.byref(d), C.byref(dl), C.byref(reg)))
        m = self.num_arcs
        sh = (max(m, 1),)
        return (np.ctypeslib.as_array(t, shape=sh)[:m].copy(), np.ctypeslib.as_array(h, shape=sh)[:m].copy(),
                np.ctypeslib.as_array(d, shape=sh)[:m].copy(), dl.value, bool(reg.value))


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
