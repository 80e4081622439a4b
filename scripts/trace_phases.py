"""Per-phase timeline of the resident kernels from the in-kernel trace marks (diagnostics)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen  # noqa: E402

arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k = 64 if arcs <= 1_000_000 else 24
inst = datagen.gen_kkt(arcs, 3, 1, "aa")
op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
op.set_mode(mode)
b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
dec = alg.lanczos_pass_one(op, b, k)
op.trace_enable(k)


def report(name, tr, marks, labels):
    tr = tr.astype(np.int64)
    steps = slice(8, k - 2)
    base = tr[:, steps, 0:1]
    rel = tr[:, steps, :] - base  # cycles since step start, per CTA
    print(f"== {name}: cycles since step start (mean over CTAs and steps | min | max over CTAs of the step mean)")
    prev = None
    for m, lab in zip(marks, labels):
        v = rel[:, :, m].mean(axis=1)
        d = "" if prev is None else f"  delta {v.mean() - prev:9.0f}"
        print(f"  mark {m:2d} {lab:28s} {v.mean():9.0f} | {v.min():9.0f} | {v.max():9.0f}{d}")
        prev = v.mean()
    step_len = (tr[:, 9:k - 2, 0] - tr[:, 8:k - 3, 0]).mean()
    print(f"  step length {step_len:.0f} cycles")
    g = tr[:, steps, -1]
    print(f"  globaltimer skew of step start across CTAs: mean {np.mean(g.max(axis=0) - g.min(axis=0)):.0f} ns")
    gl = (tr[0, 9:k - 2, -1] - tr[0, 8:k - 3, -1]).mean()
    print(f"  step length by globaltimer {gl:.0f} ns -> SM clock {step_len / gl:.3f} GHz")


alg.lanczos_pass_one(op, b, k)
t1 = op.trace_read()
report("pass 1 resident", t1, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12],
       ["step start", "gather issued+reduced", "after sync", "arcs+rows (phase A) done", "alpha sync: after bar",
        "alpha: published", "alpha: all slots seen", "alpha: after final bar", "phase B done", "beta sync: after bar",
        "beta: published", "beta: all slots seen", "beta: after final bar"])
y = np.ones(dec.steps_taken)
alg.lanczos_pass_two(op, b, dec, y)
t2 = op.trace_read()
if mode != 0 or arcs > 1_000_000:
    tt = t1.astype(np.int64)[:, 8:k - 2, 16:20]
    for nm, col in zip(("stream", "sync1", "fold(thread 0)", "sync2"), range(4)):
        v = tt[:, :, col].mean(axis=1)
        print(f"  phase B {nm:16s} {v.mean():9.0f} | {v.min():9.0f} | {v.max():9.0f}")
    sys.exit(0)
report("pass 2 resident", t2, [0, 1, 2, 3, 4, 5, 6, 7],
       ["step start", "gather issued+reduced", "after sync", "arcs+rows done", "sync: after bar", "published",
        "all slots seen", "after final bar"])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "trace.npz"), pass1=t1, pass2=t2)
