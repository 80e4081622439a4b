"""Host-side logic of the arc-partitioned multi-GPU mode on CPU: block arithmetic, vector layout, and -- with a
world_size-2 `gloo` group -- the rendezvous plus the algebra the device path relies on: the KKT product of the whole
operator equals, per rank, complete arc rows from (local arcs, replicated node segment) and node rows that are the
all-reduced partial sums E_r x_arc."""
import os
import socket

import numpy as np
import pytest

import helpers
from two_pass_lanczos_b200 import datagen, sharding


def test_arc_range_partitions_exactly():
    for m in (0, 1, 7, 1000, 500_000, 50_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.arc_range(m, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == m
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.arc_range(10, 2, 2)


def test_local_vector_roundtrip():
    m, p, world = 103, 11, 4
    v = np.arange(m + p, dtype=np.float64)
    parts = [sharding.local_vector(v, m, p, *sharding.arc_range(m, r, world)) for r in range(world)]
    assert all(len(x) == (hi - lo) + p for x, (lo, hi) in zip(parts, [sharding.arc_range(m, r, world) for r in range(world)]))
    assert np.array_equal(sharding.assemble_global(parts, m, p), v)
    with pytest.raises(ValueError):
        sharding.local_vector(v[:-1], m, p, 0, 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, q):
    import torch.distributed as dist
    import torch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inst = datagen.gen_kkt(m, 3, 5, "wc")
        p = inst.p
        # rendezvous: rank 0's 128-byte id reaches everybody (a stand-in id: NCCL itself needs a GPU)
        ident = sharding.broadcast_unique_id(dist, rank, make=lambda: bytes(range(128)))
        assert ident == bytes(range(128))
        lo, hi = sharding.arc_range(m, rank, world)
        x = helpers.seeded_b(inst.n, seed=11)
        xl = sharding.local_vector(x, m, p, lo, hi)
        t, h, d = inst.tail[lo:hi].astype(np.int64), inst.head[lo:hi].astype(np.int64), inst.d[lo:hi]
        xa, xn = xl[: hi - lo], xl[hi - lo:]
        # complete arc rows (the node segment is replicated), in the reference's CSC accumulation order: D x first, then
        # the two node columns in ascending node index
        y_arc = np.where(t < h, (d * xa + xn[t]) - xn[h], (d * xa - xn[h]) + xn[t])
        part = np.zeros(p)
        np.add.at(part, t, xa)
        np.subtract.at(part, h, xa)                          # partial node sums of this rank's arcs
        red = torch.from_numpy(part)
        dist.all_reduce(red)                                 # the per-step collective of the device path
        yl = np.concatenate([y_arc, red.numpy()])
        gathered = [None] * world
        dist.all_gather_object(gathered, yl)
        if rank == 0:
            y = sharding.assemble_global(gathered, m, p)
            y_ref = helpers.oracle_op(inst).apply(x)
            q.put((helpers.rel(y, y_ref), bool(np.array_equal(y[:m], y_ref[:m])),
                   max(float(np.abs(g[len(g) - p:] - gathered[0][len(gathered[0]) - p:]).max()) for g in gathered)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_product_matches_oracle():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, 2000, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    err, arcs_exact, replica_gap = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert arcs_exact                 # arc rows follow the reference's accumulation order exactly
    assert err < 1e-14
    assert replica_gap == 0.0         # node replicas are bit-identical after the all-reduce


class _StubOp:
    """Stands in for a sharded LinOp on a box without GPUs: records the handles `connect_fabric` hands over."""

    def __init__(self, rank, has_block=True, can_map=True):
        self.rank, self.has_block, self.can_map, self.imported, self.mode = rank, has_block, can_map, None, 0

    def fabric_export(self):
        if not self.has_block:
            raise RuntimeError("no exchange block")
        return bytes([self.rank]) * 64

    def fabric_import(self, handles):
        if not self.can_map:
            raise RuntimeError("CUDA error: peer mapping failed")
        self.imported = list(handles)

    def set_mode(self, mode):
        self.mode = mode


def _fabric_worker(rank, world, port, q, broken_rank, unmappable_rank=-1):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        op = _StubOp(rank, has_block=rank != broken_rank, can_map=rank != unmappable_rank)
        ok = sharding.connect_fabric(op, dist)
        q.put((rank, ok, op.imported, op.mode))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("broken_rank", [-1, 1])
def test_world2_gloo_fabric_rendezvous(broken_rank):
    """Every rank receives all 64-byte handles in rank order; if one rank has no exchange block, nobody connects."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_fabric_worker, args=(r, world, port, q, broken_rank)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, ok, imported, mode in got:
        if broken_rank < 0:
            assert ok and imported == [bytes([0]) * 64, bytes([1]) * 64] and mode == 0
        else:
            assert not ok and imported is None


def test_world2_gloo_fabric_falls_back_collectively_when_one_rank_cannot_map():
    """rank 1 cannot map its peer's block: both ranks leave the fused path (mode 1 = NCCL phase kernels), nobody hangs"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fabric_worker, args=(r, 2, port, q, -1, 1)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert [(ok, mode) for _, ok, _, mode in got] == [(False, 1), (False, 1)]
