"""Arc-partitioned multi-GPU mode: host-side plumbing around `tpl_op_from_kkt_sharded`.

Rank r of `world` owns the contiguous arc block `arc_range(m, r, world)` and a replica of the p node entries of every
vector; a rank-local vector is `[arc slice | node part]`.  The collectives on the data path (node-sum reduce-scatter,
node-value all-gather, alpha / beta all-reduce per Lanczos step) run inside the library: fused into one persistent kernel
per rank over peer memory once `connect_fabric` has mapped the ranks' exchange blocks, otherwise as NCCL calls between
phase kernels.  This module only does the rendezvous (`torch.distributed`, any backend) and the slicing / gathering of
vectors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, c_u32p
from .operators import LinOp


def arc_range(m: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced arc blocks: the first m % world ranks own one arc more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_vector(v_global, m: int, p: int, lo: int, hi: int):
    """Rank-local layout of a global vector [arcs (m) | nodes (p)]: its arc slice followed by all node entries."""
    v = np.asarray(v_global)
    if v.shape[0] != m + p:
        raise ValueError("vector length must be m + p")
    return np.concatenate([v[lo:hi], v[m:]])


def assemble_global(parts, m: int, p: int):
    """Inverse of `local_vector` given every rank's local vector (in rank order): arc slices concatenated, node part taken
    from rank 0 (all replicas are identical by construction; the caller may assert it)."""
    arcs = np.concatenate([np.asarray(x)[: len(x) - p] for x in parts])
    if arcs.shape[0] != m:
        raise ValueError("arc slices do not add up to m")
    return np.concatenate([arcs, np.asarray(parts[0])[len(parts[0]) - p:]])


def unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 makes it, everybody else receives it through `broadcast_unique_id`)."""
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.load().tpl_comm_unique_id(buf))
    return bytes(buf)


def broadcast_unique_id(dist, rank: int, make=unique_id) -> bytes:
    """Rendezvous over an initialised torch.distributed process group (gloo or nccl)."""
    box = [make() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def connect_fabric(op: LinOp, dist) -> bool:
    """Exchanges the CUDA IPC handles of the ranks' exchange blocks over an initialised torch.distributed group and maps
    them: the passes of `op` then run as one persistent kernel per rank with the collectives fused in (peer-memory stores
    over NVLink).  Returns False (and leaves the NCCL phase kernels in charge on EVERY rank) when a rank has no exchange
    block or cannot map a peer's (no peer access / CUDA IPC between the processes): the decision is collective, a rank never
    runs fused against one that does not."""
    world = dist.get_world_size()
    try:
        mine = op.fabric_export()
    except Exception:  # noqa: BLE001 - e.g. the tiled kernels do not fit this rank's block
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    if any(h is None for h in handles):
        return False
    try:
        op.fabric_import(handles)
        mapped = True
    except Exception:  # noqa: BLE001
        mapped = False
    flags = [None] * world
    dist.all_gather_object(flags, mapped)  # also the barrier: nobody starts a fused pass before every rank has mapped its peers
    if not all(flags):
        op.set_mode(1)  # NCCL phase kernels on every rank, also on those whose own import succeeded
        return False
    return True


def sharded_linop(m: int, p: int, tail, head, d, rank: int, world: int, nccl_id: bytes, device: int = -1, dist=None) -> LinOp:
    """LinOp over this rank's arc block of A = [[D, E^T], [E, 0]] (global tail / head / d arrays are passed; the
    library slices them).  With `dist` (an initialised torch.distributed module) and world > 1 the ranks' exchange
    blocks are connected (`connect_fabric`)."""
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    d = np.ascontiguousarray(d, dtype=np.float64)
    lo, hi = arc_range(m, rank, world)
    idbuf = (C.c_uint8 * 128).from_buffer_copy(nccl_id)
    h = C.c_void_p()
    _lib.check(_lib.load().tpl_op_from_kkt_sharded(m, p, lo, hi, tail.ctypes.data_as(c_u32p), head.ctypes.data_as(c_u32p),
                                                   d.ctypes.data_as(c_dp), len(d), device, rank, world, idbuf,
                                                   C.byref(h)))
    op = LinOp(h)
    op.arc_lo, op.arc_hi, op.m_global, op.p = lo, hi, m, p
    op.fused = bool(dist is not None and world > 1 and connect_fabric(op, dist))
    return op
