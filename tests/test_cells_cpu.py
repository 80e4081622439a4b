"""Host-side tables of the 2-D cell partition (tpl_cells_host.h) checked without a GPU: every arc in exactly one slot,
local node indices decode to the arc's own tail / head, and the node sums formed through the item / slot tables from an
integer-valued arc vector equal E w exactly (consistency code 0)."""
import ctypes as C

import numpy as np
import pytest

from two_pass_lanczos_b200 import _lib, datagen

SMEM = 232448  # opt-in shared memory per CTA on sm_100


def plan(m, p, tail, head, ctas=148, smem=SMEM):
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    stats = (C.c_uint64 * 16)()
    _lib.check(_lib.load().tpl_cells_plan(m, p, tail.ctypes.data_as(_lib.c_u32p), head.ctypes.data_as(_lib.c_u32p), ctas, smem,
                                          stats))
    keys = ["fits", "GR", "GC", "Amax", "L", "max_rows", "max_groups", "max_lines", "max_slots", "max_own", "inbox_atoms",
            "largest", "smallest", "smem", "code"]
    out = dict(zip(keys, list(stats)))
    out["conflicts_arc_rows"] = (stats[15] & 0xffffffff) / 1000.0
    out["conflicts_node_sums"] = (stats[15] >> 32) / 1000.0
    return out


@pytest.mark.parametrize("m,rho", [(1000, 1), (1000, 3), (5000, 3), (50_000, 3), (50_000, 1), (500_000, 3)])
def test_netgen_shaped_instances(m, rho):
    inst = datagen.gen_kkt(m, rho, 7, "wc")
    st = plan(inst.m, inst.p, inst.tail, inst.head)
    assert st["fits"] == 1 and st["code"] == 0, st
    assert st["GR"] * st["GC"] <= 148 and st["L"] == (inst.p + 7) // 8
    assert st["smem"] <= SMEM
    if m >= 50_000:  # balanced: the largest cell stays close to the mean
        assert st["largest"] <= 1.25 * m / (st["GR"] * st["GC"])
    if m >= 500_000:  # the placement keeps the gathers close to conflict-free (random placement: ~3 wavefronts)
        assert st["conflicts_arc_rows"] < 1.6 and st["conflicts_node_sums"] < 1.6, st


@pytest.mark.parametrize("ctas", [2, 7, 16, 132, 148])
def test_any_grid(ctas):
    inst = datagen.gen_kkt(5000, 3, 3, "aa")
    st = plan(inst.m, inst.p, inst.tail, inst.head, ctas=ctas)
    assert st["fits"] == 1 and st["code"] == 0, st


def test_self_loops_duplicates_and_isolated_nodes():
    rng = np.random.default_rng(5)
    p, m = 203, 4000
    tail = rng.integers(0, p - 20, m).astype(np.uint32)   # the last 20 nodes never appear as tails
    head = rng.integers(10, p - 10, m).astype(np.uint32)  # ... and some nodes are isolated altogether
    tail[::17] = head[::17]                               # self-loops
    tail[1::31], head[1::31] = tail[0], head[0]           # parallel arcs
    st = plan(m, p, tail, head)
    assert st["fits"] == 1 and st["code"] == 0, st


def test_unsorted_arcs_and_tiny_node_sets():
    rng = np.random.default_rng(9)
    for p in (1, 2, 5, 9, 64):
        m = 300
        tail = rng.integers(0, p, m).astype(np.uint32)
        head = rng.integers(0, p, m).astype(np.uint32)
        st = plan(m, p, tail, head)
        assert st["fits"] == 1 and st["code"] == 0, (p, st)


def test_too_large_for_shared_memory_is_reported():
    inst = datagen.gen_kkt(50_000, 3, 1, "wc")
    st = plan(inst.m, inst.p, inst.tail, inst.head, ctas=4)  # 12 500 arcs per cell: more than a thread block holds in registers
    assert st["fits"] == 0
    inst = datagen.gen_kkt(800_000, 3, 1, "wc")
    assert plan(inst.m, inst.p, inst.tail, inst.head)["fits"] == 0


def test_random_multigraphs_property():
    """hypothesis: whenever the planner says an instance fits, its tables are consistent (code 0) -- any multigraph with
    self-loops, parallel arcs, isolated nodes and unsorted tails, any grid size"""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(st.integers(1, 400), st.integers(1, 6000), st.sampled_from([1, 4, 16, 36, 100, 148]), st.integers(0, 2**31),
           st.sampled_from(["uniform", "sorted", "hub"]))
    def run(p, m, ctas, seed, shape):
        rng = np.random.default_rng(seed)
        tail = rng.integers(0, p, m)
        head = rng.integers(0, p, m)
        if shape == "sorted":
            tail = np.sort(tail)
        elif shape == "hub":
            tail[: m // 2] = tail[0]
            head[m // 2:] = head[-1]
        st_ = plan(m, p, tail, head, ctas=ctas)
        if st_["fits"]:
            assert st_["code"] == 0, (p, m, ctas, seed, shape, st_)
            assert st_["smem"] <= SMEM and st_["GR"] * st_["GC"] <= ctas

    run()
