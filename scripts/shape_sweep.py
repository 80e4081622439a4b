"""Robustness sweep: default-mode solves of odd-shaped KKT instances against the CPU oracle (diagnostics).
usage: shape_sweep.py"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import two_pass_lanczos_b200 as tpl  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def oracle_op(m, p, tail, head, d):
    j = np.arange(m, dtype=np.uint64)
    t, h = tail.astype(np.uint64), head.astype(np.uint64)
    ones = np.ones(m)
    dd = np.zeros(m)
    dd[: len(d)] = d
    return orc.SparseColMat.try_new_from_triplets(
        m + p, m + p, np.concatenate([j, m + t, m + h, j, j]), np.concatenate([j, j, j, m + t, m + h]),
        np.concatenate([dd, ones, -ones, ones, -ones]))


def run(name, m, p, seed=0, dlen=None, grouped=False, k=12):
    rng = np.random.default_rng(seed)
    tail = rng.integers(0, p, m).astype(np.uint32)
    head = rng.integers(0, p, m).astype(np.uint32)
    if grouped:
        tail = np.sort(tail)
    d = rng.uniform(1.0, 10.0, m if dlen is None else dlen)
    try:
        gop = tpl.LinOp.from_kkt(m, p, tail, head, d)
        oop = oracle_op(m, p, tail, head, d)
        x = rng.standard_normal(m + p)
        e_apply = np.linalg.norm(gop.apply(x) - oop.apply(x)) / max(np.linalg.norm(oop.apply(x)), 1e-300)
        b = rng.standard_normal(m + p)
        b /= np.linalg.norm(b)
        xg = tpl.lanczos_two_pass(gop, b, k, "exp")
        xc = orc.lanczos_two_pass(oop, b, k, npo.exp_tk_solver)
        e_x = np.linalg.norm(xg - xc) / max(np.linalg.norm(xc), 1e-300)
        x1 = tpl.lanczos(gop, b, k, "exp")
        e_1 = np.linalg.norm(x1 - xc) / max(np.linalg.norm(xc), 1e-300)
        ok = e_apply < 1e-13 and e_x < 1e-10 and e_1 < 1e-10
        print(f"{'ok  ' if ok else 'FAIL'} {name:28s} m={m:8d} p={p:8d} shape={gop.kernel_shape():8s} apply {e_apply:.1e} two-pass {e_x:.1e} one-pass {e_1:.1e}", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"EXC  {name:28s} m={m} p={p}: {type(exc).__name__}: {str(exc)[:160]}", flush=True)
        traceback.print_exc(limit=1)


run("one arc", 1, 2)
run("one arc one node", 1, 1)
run("two arcs", 2, 3)
run("31 arcs", 31, 7)
run("129 arcs", 129, 40)
run("m < p", 500, 5000)
run("no costs", 3000, 100, dlen=0)
run("cells upper edge", 144 * 4320, 1200, grouped=True)
run("just above cells", 144 * 4320 + 1, 1200, grouped=True)
run("650k", 650_000, 1400, grouped=True)
run("700k random", 700_000, 1500)
run("1.2M", 1_200_000, 2000, grouped=True)
run("blocked threshold", 1_480_000, 2200, grouped=True)
run("2^20 - 1", (1 << 20) - 1, 1800, grouped=True)
run("2^20", 1 << 20, 1800, grouped=True)
run("2^20 random order", 1 << 20, 1800)
run("1.1M p=200k", 1_100_000, 200_000)
run("1.1M p=40k", 1_100_000, 40_000)
run("3M p=3", 3_000_000, 3, grouped=True)
run("3M p=120k", 3_000_000, 120_000, grouped=True)
run("3M p=500k", 3_000_000, 500_000)
run("6M p=30", 6_000_000, 30, grouped=True, k=8)
