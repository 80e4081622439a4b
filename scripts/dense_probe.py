"""Dense operator: two-pass vs one-pass at fixed n over k (mirrors src/bin/dense_tradeoff.rs) and the achieved bandwidth of the
fused matvec + recurrence kernels (8 n^2 algorithmic bytes per step).  usage: dense_probe.py [n] [k ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ks = [int(a) for a in sys.argv[2:]] or [50, 200]
g = np.random.default_rng(1).standard_normal((n, n))
a = (g + g.T) / (2 * np.sqrt(n)) + 3.0 * np.eye(n)
op = tpl.LinOp.from_dense(a)
b = np.random.default_rng(2).random(n)
print("n,k,variant,wall_ms,pass1_ms,pass2_ms,gemv_ms,pass1_GBs,frac_of_6546.6,rel_diff_vs_two_pass")
for k in ks:
    x2 = None
    for variant in ("two-pass", "one-pass"):
        for rep in range(2):
            t = time.time()
            x = tpl.lanczos_two_pass(op, b, k, "inv") if variant == "two-pass" else tpl.lanczos(op, b, k, "inv")
            wall = time.time() - t
            tm = op.last_timing()
        gbs = k * (8.0 * n * n + 48.0 * n) / (tm["pass_one_ms"] * 1e-3) / 1e9
        diff = 0.0 if x2 is None else float(np.linalg.norm(x - x2) / np.linalg.norm(x2))
        x2 = x if x2 is None else x2
        print(f"{n},{k},{variant},{wall*1e3:.2f},{tm['pass_one_ms']:.3f},{tm['pass_two_ms'] if variant == 'two-pass' else 0.0:.3f},"
              f"{tm['gemv_ms'] if variant == 'one-pass' else 0.0:.3f},{gbs:.0f},{gbs/6546.6:.3f},{diff:.2e}", flush=True)
