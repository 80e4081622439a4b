"""Blocked streaming layout (tpl_blocks_host.h), built and checked on the host through tpl_blocks_plan: gidx is a bijection
between the non-padding cell-order positions and the arcs, every packed tail/head word decodes to its arc inside the cell's
node blocks (with the CSC-order and self-loop flags), cells are sorted by (tail, arc index) and padded to whole 128-arc stages,
every non-loop arc is once on its local tail and once on its local head in the lists of its tile, the list slices have equal
length with correct new-node flags, slot fields and chain depths, a host emulation of the kernel's list walk (running sums,
slot flushes, scratch shares added depth by depth) over exact integer values reproduces every node sum of every tile
(check_cell_lists), and the layout does not depend on the number of host threads."""
import ctypes as C

import numpy as np
import pytest

from two_pass_lanczos_b200 import _lib, datagen
from two_pass_lanczos_b200._lib import c_dp, c_u32p

SMEM = 232_448  # opt-in shared memory per CTA on sm_100a
KEYS = ("fits", "GR", "GC", "PT", "PH", "Mpad", "T", "ntile", "rings", "cell_max", "cell_min", "entries", "pieces", "hash",
        "code", "smem2")


def plan(m, p, tail, head, d=None, ctas=148, smem=SMEM, threads=1):
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    d = np.zeros(0) if d is None else np.ascontiguousarray(d, dtype=np.float64)
    st = (C.c_uint64 * 16)()
    _lib.check(_lib.load().tpl_blocks_plan(m, p, tail.ctypes.data_as(c_u32p), head.ctypes.data_as(c_u32p),
                                           d.ctypes.data_as(c_dp) if len(d) else None, len(d), ctas, smem, threads, st))
    return dict(zip(KEYS, st))


@pytest.mark.parametrize("m,flavour", [(40_000, "wc"), (700_000, "aa"), (2_000_000, "wc")])
def test_layout_is_consistent_on_netgen_shaped_instances(m, flavour):
    inst = datagen.gen_kkt(m, 3, 3, flavour)
    st = plan(inst.m, inst.p, inst.tail, inst.head, inst.d)
    assert st["fits"] == 1 and st["code"] == 0
    assert (st["GR"], st["GC"]) == (12, 12)
    assert st["Mpad"] % 128 == 0 and inst.m <= st["Mpad"] <= inst.m + 144 * 127
    assert st["T"] in (1024, 2048, 3072, 4096) and st["smem2"] <= SMEM - 3072
    # node blocks balanced by degree: no cell is far above the mean
    assert st["cell_max"] <= 1.25 * inst.m / 144 + 128
    # only the two node blocks of a cell are staged: a small fraction of p
    assert st["PT"] + st["PH"] <= 0.25 * inst.p + 4
    r1, r2, r2v = st["rings"] & 0xff, (st["rings"] >> 8) & 0xff, (st["rings"] >> 16) & 0xff
    assert 2 <= r1 <= 4 and 2 <= r2 <= 4 and 2 <= r2v <= 3


def test_layout_at_the_largest_single_gpu_size_fits_with_a_deep_ring():
    """50M arcs: p = 11 547; the two node blocks of a cell hold ~2-3 k values (the last tail block also takes the sinks, which
    have no out-arcs) -> three ring slots in pass 2 and double-buffered list blocks still fit (the tiled kernels kept two
    p-long arrays, 184 KB, in shared memory).  Only the sizing is checked here (no 50M-arc build on the CPU box): blocks_fit is exercised
    through a 2M-arc instance with the node count of the 50M one."""
    rng = np.random.default_rng(5)
    m, p = 2_000_000, 11_547
    tail = np.sort(rng.integers(0, p - 1200, m))
    head = rng.integers(100, p, m)
    st = plan(m, p, tail, head)
    assert st["fits"] == 1 and st["code"] == 0
    # (this graph has no same-tail runs of 32 arcs inside a cell, so every arc costs two list entries: the worst case for the
    # list buffers; a netgen-shaped 50M-arc instance gets 2048-arc tiles)
    assert st["T"] >= 1024 and ((st["rings"] >> 8) & 0xff) >= 3


@pytest.mark.parametrize("ctas", [13, 148])
def test_layout_on_irregular_graphs(ctas):
    """random multigraph with self-loops, parallel arcs, unsorted tails, isolated nodes, a hub with a long same-tail run, a short D"""
    rng = np.random.default_rng(ctas)
    m, p = 150_000, 1300
    tail = rng.integers(0, p - 100, m)
    head = rng.integers(50, p, m)
    tail[1000:9000] = 7
    head[2000:2100] = 7
    tail[50_000:50_040] = head[50_000:50_040]
    d = rng.random(m // 2)  # the loader's short-D quirk: rows beyond d_len have no diagonal entry
    st = plan(m, p, tail, head, d, ctas=ctas)
    assert st["fits"] == 1 and st["code"] == 0
    assert st["GR"] * st["GC"] <= ctas
    assert st["pieces"] > 0


def test_layout_does_not_depend_on_the_host_thread_count():
    inst = datagen.gen_kkt(1_200_000, 3, 9, "wc")
    a = plan(inst.m, inst.p, inst.tail, inst.head, inst.d, threads=1)
    for threads in (3, 16):
        b = plan(inst.m, inst.p, inst.tail, inst.head, inst.d, threads=threads)
        assert b == a


def test_layout_degenerate_inputs():
    st = plan(40_000, 1, np.zeros(40_000), np.zeros(40_000))  # one node: every arc is a self-loop
    assert st["fits"] == 1 and st["code"] == 0 and st["pieces"] == 0
    rng = np.random.default_rng(1)
    st = plan(33_000, 40_000, rng.integers(0, 40_000, 33_000), rng.integers(0, 40_000, 33_000))  # more nodes than arcs
    assert st["fits"] == 1 and st["code"] == 0
    st = plan(100_000, 300_000, rng.integers(0, 300_000, 100_000), rng.integers(0, 300_000, 100_000), smem=60_000)
    assert st["fits"] == 0  # two node blocks + tile buffers + ring do not fit: the caller keeps the other shapes


def test_layout_of_a_shard_ignores_nodes_without_local_arcs():
    """A rank of a sharded operator holds a contiguous slice of the arcs but all p nodes: the tails of the other ranks' arcs (and
    sinks / sources) have no arcs on that side and must not take shared memory -- blocks are cut over the ACTIVE nodes only."""
    inst = datagen.gen_kkt(1_600_000, 3, 4, "wc")
    lo, hi = inst.m // 4, inst.m // 2  # rank 1 of 4
    st = plan(hi - lo, inst.p, inst.tail[lo:hi], inst.head[lo:hi], inst.d[lo:hi])
    assert st["fits"] == 1 and st["code"] == 0
    active_tails = len(np.unique(inst.tail[lo:hi]))
    assert active_tails < 0.4 * inst.p
    assert st["PT"] <= active_tails / 12 + 8          # not p / 12 + (nodes without out-arcs)
    assert st["PH"] <= inst.p / 12 + 8


@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
def test_the_checker_catches_deliberate_defects(kind, monkeypatch):
    """The checker is what vouches for the device-built tables (tpl_op_layout_check): it must reject a layout with ONE
    defect in one tile -- a flipped sign, a sum flushed to the wrong slot, a wrong chain depth, two exchanged entries,
    a wrong slot for a slice's last node (TPL_PLAN_CORRUPT in tpl_blocks_plan)."""
    inst = datagen.gen_kkt(40_000, 3, 3, "wc")
    assert plan(inst.m, inst.p, inst.tail, inst.head, inst.d)["code"] == 0
    monkeypatch.setenv("TPL_PLAN_CORRUPT", str(kind))
    assert plan(inst.m, inst.p, inst.tail, inst.head, inst.d)["code"] != 0
