"""Timing of the cell kernels at one size, with the per-phase / per-warp trace (diagnostics).
usage: cells_exp.py [arcs] [k] [--trace] [--grid] [--skew]
(the trace itself costs ~2 000 cycles per step, most of them on warp 0, whose thread 0 writes the marks)"""
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
print(f"{op.kernel_shape()} m={arcs} k={K}: pass1 {best[0]:.3f} ms pass2 {best[1]:.3f} ms "
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


def report(name, tr, labels, extra):
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
    print(f"  step length {step_len:.0f} cycles; node sums done + barrier at {relc[:, :, extra].mean():.0f}")
    if extra == 10:
        print(f"  phase A: scalars known at {relc[:, :, 11].mean():.0f}, arc rows done (thread 0) at {relc[:, :, 12].mean():.0f}")
    for base, lab in ((32, "polls done, by warp"), (48, "sums pushed, by warp")):
        v = relc[:, :, base:base + 15].mean(axis=(0, 1))
        mx = relc[:, :, base:base + 15].max(axis=2).mean()
        print(f"  {lab:22s} " + " ".join(f"{x:6.0f}" for x in v) + f"   mean of the per-step max {mx:.0f}")


if "--grid" in sys.argv:
    for name, tr, mk in (("pass 1 after-sync (mark 2)", t1, 2), ("pass 1 inbox warp 5 polls done", t1, 37),
                         ("pass 2 inbox warp 0 polls done", t2, 32)):
        v = (tr[:144, 8:k - 3, mk] - tr[:144, 8:k - 3, 0]).mean(axis=1)
        print(f"== {name}: per-CTA mean, 12 x 12 (row = tail block)")
        for a in range(12):
            print("   " + " ".join(f"{x:6.0f}" for x in v[a * 12:(a + 1) * 12]))
if "--skew" in sys.argv:
    for name, tr, push_mark, poll_mark in (("pass 1", t1, 9, 37), ("pass 2", t2, 5, 32)):
        tr = tr[:144]
        g0 = tr[:, :, 63].astype(np.float64)                       # ns, globaltimer at step start (same clock on every SM)
        ghz = 1.965
        push = g0 + (tr[:, :, push_mark] - tr[:, :, 0]) / ghz      # ns
        poll = g0 + (tr[:, :, poll_mark] - tr[:, :, 0]) / ghz
        st = slice(8, k - 4)
        start_skew = (g0[:, st].max(axis=0) - g0[:, st].min(axis=0)).mean()
        push_spread = (push[:, st].max(axis=0) - push[:, st].min(axis=0)).mean()
        lat_after_last = (poll[:, 9:k - 3] - push[:, st].max(axis=0)[None, :]).mean()
        lat_after_mean = (poll[:, 9:k - 3] - push[:, st].mean(axis=0)[None, :]).mean()
        print(f"== {name}: step-start skew across CTAs {start_skew:.0f} ns, push spread {push_spread:.0f} ns, inbox poll done "
              f"{lat_after_last:.0f} ns after the LAST push of the grid ({lat_after_mean:.0f} ns after the mean push)")
report("pass 1 cells", t1, ["step start", "polls done (thread 0)", "after sync", "phase A done", "alpha published",
                            "alpha polled", "alpha known", "phase B done", "beta published", "sums pushed"], 10)
report("pass 2 cells", t2, ["step start", "polls done (thread 0)", "after sync", "rows done", "after sync", "sums pushed"], 6)
