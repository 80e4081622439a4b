"""Per-phase timeline of the blocked streaming kernels from the in-kernel trace marks (diagnostics).
usage: trace_blocked.py [arcs] [k]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen  # noqa: E402

arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 24
inst = datagen.gen_kkt(arcs, 3, 1, "aa")
op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
op.set_mode(5)
print(op.kernel_shape(), "arcs", arcs, "p", inst.p)
b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
dec = alg.lanczos_pass_one(op, b, k)
op.trace_enable(k)


def report(name, tr, marks, labels, extra):
    tr = tr.astype(np.int64)
    G = 144
    tr = tr[:G]
    steps = slice(4, k - 3)
    rel = tr[:, steps, :] - tr[:, steps, 0:1]
    print(f"== {name}: cycles since step start (mean over CTAs and steps | min | max over CTAs of the step mean)")
    prev = None
    for m, lab in zip(marks, labels):
        v = rel[:, :, m].mean(axis=1)
        d = "" if prev is None else f"  delta {v.mean() - prev:9.0f}"
        print(f"  mark {m:2d} {lab:34s} {v.mean():9.0f} | {v.min():9.0f} | {v.max():9.0f}{d}")
        prev = v.mean()
    step_len = (tr[:, 5:k - 3, 0] - tr[:, 4:k - 4, 0]).mean()
    print(f"  step length {step_len:.0f} cycles")
    for m, lab in extra:
        v = tr[:, steps, m].mean(axis=1)
        print(f"  sweep: {lab:40s} {v.mean():9.0f} | {v.min():9.0f} | {v.max():9.0f}")


alg.lanczos_pass_one(op, b, k)
t1 = op.trace_read()
SYNC = ["sync: after bar", "sync: published", "sync: all slots seen", "sync: after final bar"]
EXTRA = [(16, "compute warp 0 waits for bulk copies"), (17, "compute warp 0 waits for the fold"), (18, "compute warp 0 whole sweep"),
         (23, "compute warp 0 consuming stages"), (24, "compute warp 0 last stage consumed at"), (19, "fold thread 0 waits for a full tile"), (21, "fold thread 0 waits for list blocks"),
         (22, "fold thread 0 folding"), (25, "fold thread 0 preamble done at"), (26, "fold thread 0 first tile full at"), (20, "fold thread 0 whole sweep")]
report("pass 1 blocked", t1, list(range(15)),
       ["step start", "nodes staged", "owner rows A", "phase A arcs done"] + ["alpha " + s for s in SYNC] +
       ["owner rows B + zero acc", "phase B sweep done", "partials published"] + ["beta " + s for s in SYNC], EXTRA)
y = np.ones(dec.steps_taken)
alg.lanczos_pass_two(op, b, dec, y)
t2 = op.trace_read()
report("pass 2 blocked", t2, list(range(8)), ["step start", "staged + owner rows", "sweep done", "partials published"] + SYNC, EXTRA)
