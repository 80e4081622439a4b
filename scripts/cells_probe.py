"""Bring-up / timing probe of the cell kernels (diagnostics, not a test): parity against the chunk-resident kernels and the
oracle-free invariants, timing at the headline size, per-phase trace."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen  # noqa: E402

arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 500


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


for m in (1000, 50_000, arcs):
    inst = datagen.gen_kkt(m, 3, 1, "aa")
    ops = {}
    for name, mode in (("cells", 0), ("chunks", 4)):
        ops[name] = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
        ops[name].set_mode(mode)
    b = ops["cells"].apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    k = 60
    dc = alg.lanczos_pass_one(ops["cells"], b, k)
    dr = alg.lanczos_pass_one(ops["chunks"], b, k)
    print(f"m={m} steps {dc.steps_taken}/{dr.steps_taken} bnorm {dc.b_norm - dr.b_norm:.3e} "
          f"alpha {np.max(np.abs(dc.alphas - dr.alphas)) / np.abs(dr.alphas).max():.3e} "
          f"beta {np.max(np.abs(dc.betas - dr.betas)) / np.abs(dr.betas).max():.3e}", flush=True)
    if m <= 50_000:
        so = alg.lanczos_standard(ops["cells"], b, k)
        y = 0.1 * (np.arange(dc.steps_taken) + 1)
        p2 = alg.lanczos_pass_two_with_basis(ops["cells"], b, dc, y)
        sr = alg.lanczos_standard(ops["chunks"], b, k)
        print(f"   drift {np.abs(so.v_k - p2.v_k).max():.3e} one-pass coef equal {np.array_equal(so.decomposition.alphas, dc.alphas)} "
              f"basis vs chunks {np.abs(so.v_k[:, :20] - sr.v_k[:, :20]).max():.3e} x vs Vy {rel(p2.x_k, so.v_k @ y):.3e}", flush=True)
    for name in ("cells", "chunks"):
        for rep in range(3):
            t = time.time()
            x = tpl.lanczos_two_pass(ops[name], b, K, "inv")
            wall = time.time() - t
            tm = ops[name].last_timing()
        res = np.linalg.norm(ops[name].apply(x) - b) / np.linalg.norm(b)
        print(f"   {name:7s} k={K} wall {wall * 1e3:.2f} ms pass1 {tm['pass_one_ms']:.3f} ms pass2 {tm['pass_two_ms']:.3f} ms "
              f"residual {res:.3e}", flush=True)
        if name == "cells":
            xc = x
    print(f"   x cells vs chunks {rel(xc, x):.3e}", flush=True)

# per-phase trace of the cell kernels at the last size
op = ops["cells"]
k = 64
dec = alg.lanczos_pass_one(op, b, k)
op.trace_enable(k)
alg.lanczos_pass_one(op, b, k)
t1 = op.trace_read().astype(np.int64)
alg.lanczos_pass_two(op, b, dec, np.ones(dec.steps_taken))
t2 = op.trace_read().astype(np.int64)


def report(name, tr, labels):
    G = 144
    tr = tr[:G]
    steps = slice(8, k - 3)
    relc = tr[:, steps, :] - tr[:, steps, 0:1]
    print(f"== {name}: cycles since step start (mean | min | max over CTAs of the per-CTA step mean)")
    prev = 0.0
    for mk, lab in enumerate(labels):
        v = relc[:, :, mk].mean(axis=1)
        print(f"  mark {mk:2d} {lab:34s} {v.mean():8.0f} | {v.min():8.0f} | {v.max():8.0f}  delta {v.mean() - prev:8.0f}")
        prev = v.mean()
    step_len = (tr[:, 9:k - 3, 0] - tr[:, 8:k - 4, 0]).mean()
    gl = (tr[0, 9:k - 3, -1] - tr[0, 8:k - 4, -1]).mean()
    print(f"  step length {step_len:.0f} cycles = {gl:.0f} ns")


report("pass 1 cells", t1, ["step start", "polls done (this thread)", "after sync", "phase A done", "alpha published",
                            "alpha polled", "alpha known", "phase B done", "beta published", "sums pushed"])
report("pass 2 cells", t2, ["step start", "polls done (this thread)", "after sync", "rows done", "after sync", "sums pushed"])
