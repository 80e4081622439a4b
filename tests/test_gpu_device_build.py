"""Operator tables built ON THE DEVICE (tpl_build.cuh: node lists, blocked streaming layout) against the host builders, which
stay as the checkers: the device tables are downloaded and (a) pass the host consistency check of the blocked layout,
(b) hash to the same value as the host-built layout (word-for-word identical lists, packed words and arc order), (c) carry the
same node lists; and solves on a device-built handle are bit-identical to solves on a host-built one.
Reference: src/utils/data_loader.rs:211-259 (the host-side assembly this replaces for large instances)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers  # noqa: F401
import two_pass_lanczos_b200 as tpl
from two_pass_lanczos_b200 import _lib, datagen

pytestmark = pytest.mark.gpu


def _instances():
    rng = np.random.default_rng(5)
    out = {}
    inst = datagen.gen_kkt(60_000, 3, 7, "wc")
    out["wc60k"] = (inst.m, inst.p, inst.tail, inst.head, inst.d)
    inst = datagen.gen_kkt(300_000, 3, 1, "aa")
    perm = rng.permutation(inst.m)  # arcs in random order: the sort by tail has work to do
    out["aa300k_shuffled"] = (inst.m, inst.p, inst.tail[perm], inst.head[perm], inst.d[perm])
    # irregular: self-loops, parallel arcs, a hub node, isolated nodes, a short d
    m, p = 200_000, 900
    tail = rng.integers(0, p - 50, m).astype(np.uint32)
    head = rng.integers(0, p - 50, m).astype(np.uint32)
    tail[:30_000] = 7  # hub: a node that straddles many list slices
    head[30_000:36_000] = tail[30_000:36_000]  # self-loops
    tail[40_000:41_000], head[40_000:41_000] = 11, 12  # parallel arcs
    out["irregular"] = (m, p, tail, head, rng.uniform(0.5, 2.0, m - 1000))
    return out


def _host_plan(m, p, tail, head, d, ctas, smem):
    st = (C.c_uint64 * 16)()
    c_u32p, c_dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    t = np.ascontiguousarray(tail, dtype=np.uint32)
    h = np.ascontiguousarray(head, dtype=np.uint32)
    dd = np.ascontiguousarray(d, dtype=np.float64)
    _lib.check(_lib.load().tpl_blocks_plan(m, p, t.ctypes.data_as(c_u32p), h.ctypes.data_as(c_u32p), dd.ctypes.data_as(c_dp),
                                           len(dd), ctas, smem, 0, st))
    return [int(x) for x in st]


def _build(env, m, p, tail, head, d):
    for k in ("TPL_HOST_BUILD", "TPL_DEVICE_BUILD"):
        os.environ.pop(k, None)
    if env:
        os.environ[env] = "1"
    try:
        return tpl.LinOp.from_kkt(m, p, tail, head, d)
    finally:
        os.environ.pop(env, None) if env else None


@pytest.mark.parametrize("name", ["wc60k", "aa300k_shuffled", "irregular"])
def test_device_tables_equal_host_tables(name):
    import torch
    m, p, tail, head, d = _instances()[name]
    props = torch.cuda.get_device_properties(0)
    dev = _build("TPL_DEVICE_BUILD", m, p, tail, head, d)
    got = dev.layout_check(tail, head, d)
    assert got["device_built"] == 1 and got["blocked"] == 1
    assert got["check"] == 0, got
    assert got["node_list_mismatches"] == 0 and got["node_list_words"] > 0
    plan = _host_plan(m, p, tail, head, d, props.multi_processor_count, props.shared_memory_per_block_optin)
    assert plan[0] == 1 and plan[14] == 0
    assert got["tile_arcs"] == plan[6] and got["list_words"] == plan[11]
    assert got["hash"] == plan[13]
    host = _build("TPL_HOST_BUILD", m, p, tail, head, d)
    hgot = host.layout_check(tail, head, d)
    assert hgot["device_built"] == 0 and hgot["check"] == 0 and hgot["hash"] == got["hash"]
    # the same solve on both handles, blocked kernels and generic kernels
    n = m + p
    b = dev.apply(np.full(n, 1.0 / np.sqrt(n)))
    assert np.array_equal(b, host.apply(np.full(n, 1.0 / np.sqrt(n))))
    for mode in (5, 1):
        dev.set_mode(mode)
        host.set_mode(mode)
        assert np.array_equal(tpl.lanczos_two_pass(dev, b, 40, "inv"), tpl.lanczos_two_pass(host, b, 40, "inv"))


def test_large_instances_build_on_the_device_by_default():
    inst = datagen.gen_kkt(2_000_000, 3, 7, "wc")
    op = _build(None, inst.m, inst.p, inst.tail, inst.head, inst.d)
    got = op.layout_check(inst.tail, inst.head, inst.d)
    assert got["device_built"] == 1 and got["blocked"] == 1 and got["check"] == 0 and got["node_list_mismatches"] == 0
    assert op.kernel_shape() == "blocked"
    small = datagen.gen_kkt(60_000, 3, 7, "wc")
    assert _build(None, small.m, small.p, small.tail, small.head, small.d).layout_check(small.tail, small.head, small.d)["device_built"] == 0


@pytest.mark.parametrize("case", ["all_loops", "many_nodes", "two_nodes"])
def test_degenerate_large_instances_build_like_the_host(case):
    """Shapes the blocked layout cannot or need not hold: the device path must end where the host path ends."""
    rng = np.random.default_rng(11)
    m = (1 << 20) + 77
    if case == "all_loops":  # every arc a self-loop: no incidence entries at all, A = diag(D, 0)
        p = 3
        tail = rng.integers(0, p, m).astype(np.uint32)
        head = tail.copy()
    elif case == "many_nodes":  # more nodes than a block holds (local ids are 15 bits): gather kernels
        p = 3_000_000
        tail = rng.integers(0, p, m).astype(np.uint32)
        head = rng.integers(0, p, m).astype(np.uint32)
    else:  # one tail and one head node: a single run per stage, the hub straddles every slice
        p = 2
        tail = np.zeros(m, dtype=np.uint32)
        head = np.ones(m, dtype=np.uint32)
    d = rng.uniform(1.0, 2.0, m)
    dev = _build(None, m, p, tail, head, d)
    host = _build("TPL_HOST_BUILD", m, p, tail, head, d)
    got, want = dev.layout_check(tail, head, d), host.layout_check(tail, head, d)
    assert got["device_built"] == 1 and want["device_built"] == 0
    assert got["blocked"] == want["blocked"] and got["check"] == 0 and want["check"] == 0
    assert got["hash"] == want["hash"] and got["node_list_mismatches"] == 0
    assert dev.kernel_shape() == host.kernel_shape()
    n = m + p
    b = rng.standard_normal(n)
    assert np.array_equal(dev.apply(b), host.apply(b))
    assert np.array_equal(tpl.lanczos_two_pass(dev, b, 12, "inv"), tpl.lanczos_two_pass(host, b, 12, "inv"))


@pytest.mark.parametrize("builder", ["TPL_HOST_BUILD", "TPL_DEVICE_BUILD"])
def test_sparse_graph_with_a_million_nodes_matches_the_oracle(builder):
    """More node rows than a CTA's shared memory holds segment sums for (they go through HBM scratch, tpl_kernels.cuh
    seg_put / seg_get) and more nodes than the blocked layout's 15-bit local ids: the gather kernels, against the oracle."""
    from oracle import np_oracle as npo
    from oracle import oracle as orc
    rng = np.random.default_rng(23)
    p, m = 1_000_000, 1_500_000
    tail = rng.integers(0, p, m).astype(np.uint32)
    head = rng.integers(0, p, m).astype(np.uint32)
    d = rng.uniform(1.0, 10.0, m)
    j = np.arange(m, dtype=np.uint64)
    t, h = tail.astype(np.uint64), head.astype(np.uint64)
    ones = np.ones(m)
    oop = orc.SparseColMat.try_new_from_triplets(
        m + p, m + p, np.concatenate([j, m + t, m + h, j, j]), np.concatenate([j, j, j, m + t, m + h]),
        np.concatenate([d, ones, -ones, ones, -ones]))
    gop = _build(builder, m, p, tail, head, d)
    assert gop.kernel_shape() == "gather"
    x = rng.standard_normal(m + p)
    assert helpers.rel(gop.apply(x), oop.apply(x)) < 1e-14
    b = helpers.seeded_b(m + p)
    x_gpu = tpl.lanczos_two_pass(gop, b / np.linalg.norm(b), 25, "exp")
    x_cpu = orc.lanczos_two_pass(oop, b / np.linalg.norm(b), 25, npo.exp_tk_solver)
    assert helpers.rel(x_gpu, x_cpu) <= 1e-10
