"""Timing of the cell kernels at one size, with the per-phase / per-warp trace (diagnostics).
usage: cells_exp.py [arcs] [k] ; TPL_CELL_FLAGS selects experiment switches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen  # noqa: E402

arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 500
inst = datagen.gen_kkt(arcs, 3, 1, "aa")
op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
best = (1e9, 1e9)
for rep in range(4):
    x = tpl.lanczos_two_pass(op, b, K, "inv")
    tm = op.last_timing()
    best = min(best, (tm["pass_one_ms"], tm["pass_two_ms"]))
res = np.linalg.norm(op.apply(x) - b) / np.linalg.norm(b)
print(f"flags={os.environ.get('TPL_CELL_FLAGS', '0')} m={arcs} k={K}: pass1 {best[0]:.3f} ms pass2 {best[1]:.3f} ms "
      f"total {best[0] + best[1]:.3f} ms residual {res:.3e}", flush=True)
if "--trace" not in sys.argv:
    sys.exit(0)
k = 64
dec = alg.lanczos_pass_one(op, b, k)
op.trace_enable(k)
alg.lanczos_pass_one(op, b, k)
t1 = op.trace_read().astype(np.int64)
alg.lanczos_pass_two(op, b, dec, np.ones(dec.steps_taken))
t2 = op.trace_read().astype(np.int64)


def report(name, tr, labels):
    tr = tr[:144]
    steps = slice(8, k - 3)
    relc = tr[:, steps, :] - tr[:, steps, 0:1]
    print(f"== {name}: cycles since step start (mean | min | max over CTAs of the per-CTA step mean)")
    prev = 0.0
    for mk, lab in enumerate(labels):
        v = relc[:, :, mk].mean(axis=1)
        print(f"  mark {mk:2d} {lab:28s} {v.mean():8.0f} | {v.min():8.0f} | {v.max():8.0f}  delta {v.mean() - prev:8.0f}")
        prev = v.mean()
    step_len = (tr[:, 9:k - 3, 0] - tr[:, 8:k - 4, 0]).mean()
    print(f"  step length {step_len:.0f} cycles")
    for base, lab in ((32, "polls done, by warp"), (48, "sums pushed, by warp")):
        v = relc[:, :, base:base + 15].mean(axis=(0, 1))
        mx = relc[:, :, base:base + 15].max(axis=2).mean()
        print(f"  {lab:22s} " + " ".join(f"{x:6.0f}" for x in v) + f"   mean of the per-step max {mx:.0f}")


report("pass 1 cells", t1, ["step start", "polls done (thread 0)", "after sync", "phase A done", "alpha published",
                            "alpha polled", "alpha known", "phase B done", "beta published", "sums pushed"])
report("pass 2 cells", t2, ["step start", "polls done (thread 0)", "after sync", "rows done", "after sync", "sums pushed"])
