"""Tile entry lists of the tiled streaming kernels (tpl_tiles_host.h), built and checked on the host through
tpl_tiles_plan: every non-loop arc of a tile is on its head node once (sign set) and on its tail node once (directly or
inside one same-tail piece), a node's entries of a tile sit in one thread's slice (the fold has no atomics), and the lists
do not depend on how many host threads built them."""
import ctypes as C

import numpy as np
import pytest

from two_pass_lanczos_b200 import _lib, datagen
from two_pass_lanczos_b200._lib import c_u32p
from two_pass_lanczos_b200.error import DataLoaderError, LanczosError


def plan(m, p, tail, head, ctas=148, tile=8192, threads=1):
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    st = (C.c_uint64 * 8)()
    _lib.check(_lib.load().tpl_tiles_plan(m, p, tail.ctypes.data_as(c_u32p), head.ctypes.data_as(c_u32p), ctas, tile,
                                          threads, st))
    return dict(zip(("code", "ntile", "entries", "pieces", "pads", "longest", "hash", "fold_threads"), st))


@pytest.mark.parametrize("m,flavour", [(3_000, "wc"), (50_000, "aa"), (700_000, "wc"), (2_000_000, "aa")])
def test_tile_lists_are_consistent_on_netgen_shaped_instances(m, flavour):
    inst = datagen.gen_kkt(m, 3, 3, flavour)
    st = plan(inst.m, inst.p, inst.tail, inst.head)
    assert st["code"] == 0
    arcs_per_cta = -(-inst.m // 148)
    assert st["ntile"] == max(1, -(-arcs_per_cta // 8192))
    assert st["entries"] % st["fold_threads"] == 0
    # sorted tails: most of the tail side is summed as pieces, so fewer than 2 entries per arc remain
    assert st["entries"] - st["pads"] <= 2 * inst.m
    if m >= 700_000:
        assert st["pieces"] > 0 and st["entries"] - st["pads"] < 1.5 * inst.m
        assert st["pads"] < (0.35 if m < 1_000_000 else 0.25) * st["entries"]  # slices are balanced


@pytest.mark.parametrize("tile", [512, 2048, 8192])
def test_tile_lists_on_irregular_graphs(tile):
    """random multigraph with self-loops, parallel arcs, unsorted tails, one hub node with a long same-tail run that is
    cut into pieces at tile boundaries"""
    rng = np.random.default_rng(tile)
    m, p = 120_000, 900
    tail = rng.integers(0, p, m)
    head = rng.integers(0, p, m)
    tail[1000:9000] = 7  # a run of 8000 arcs with the same tail -> pieces of <= 256, more than one tile at tile = 512
    head[2000:2100] = 7  # self-loops inside the run break it
    tail[50_000:50_040] = head[50_000:50_040]
    st = plan(m, p, tail, head, ctas=13, tile=tile)
    assert st["code"] == 0
    assert st["pieces"] >= 8000 // 256
    loops = int(np.sum(tail == head))
    assert st["entries"] - st["pads"] <= 2 * (m - loops)


def test_tile_lists_do_not_depend_on_the_host_thread_count():
    inst = datagen.gen_kkt(300_000, 3, 9, "wc")
    a = plan(inst.m, inst.p, inst.tail, inst.head, threads=1)
    for threads in (2, 5, 16):
        assert plan(inst.m, inst.p, inst.tail, inst.head, threads=threads) == a


def test_tile_plan_edge_cases_and_errors():
    assert plan(0, 4, [], [], ctas=4)["code"] == 0                       # no arcs
    assert plan(3, 2, [0, 0, 1], [0, 0, 1], ctas=4)["code"] == 0         # only self-loops: nothing to sum
    st = plan(5, 3, [0, 1, 2, 0, 1], [1, 2, 0, 2, 0], ctas=148)          # fewer arcs than CTAs
    assert st["code"] == 0 and st["entries"] - st["pads"] == 10
    with pytest.raises(DataLoaderError):
        plan(2, 2, [0, 2], [1, 1])                                       # node id out of range
    with pytest.raises(LanczosError):
        plan(2, 2, [0, 1], [1, 0], tile=16384)                           # tile + piece slots exceed the 14-bit index


def test_tile_lists_on_random_multigraphs_property():
    """hypothesis: any multigraph (self-loops, parallel arcs, isolated nodes, unsorted tails), any grid, any tile size"""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(st.integers(1, 300), st.integers(0, 4000), st.integers(1, 40), st.sampled_from([64, 256, 1024, 4096]),
           st.integers(0, 2**31), st.sampled_from(["uniform", "sorted", "hub"]))
    def run(p, m, ctas, tile, seed, shape):
        rng = np.random.default_rng(seed)
        tail = rng.integers(0, p, m)
        head = rng.integers(0, p, m)
        if shape == "sorted":
            tail = np.sort(tail)
        elif shape == "hub" and m:
            tail[: m // 2] = tail[0]  # one long same-tail run -> pieces, cut at tile and piece boundaries
        st_ = plan(m, p, tail, head, ctas=ctas, tile=tile, threads=1 + seed % 3)
        assert st_["code"] == 0, (p, m, ctas, tile, seed, shape, st_)
        loops = int(np.sum(tail == head))
        assert st_["entries"] - st_["pads"] <= 2 * (m - loops)

    run()
