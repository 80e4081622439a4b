"""Streaming-regime probe: two-pass solve at large arc counts, per-pass time and algorithmic bandwidth."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl
from two_pass_lanczos_b200 import datagen

k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sizes = [int(a) for a in sys.argv[2:]] or [5_000_000, 20_000_000, 50_000_000]
for m in sizes:
    t = time.time(); inst = datagen.gen_kkt(m, 3, 1, "aa"); tg = time.time() - t
    t = time.time(); op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d); tb = time.time() - t
    n = inst.n
    b = op.apply(np.full(n, 1.0 / np.sqrt(n)))
    for fmt_mode in (0,):
        for rep in range(3):
            x = tpl.lanczos_two_pass(op, b, k, "inv")
            tm = op.last_timing()
        bm = op.matrix_bytes()
        a1 = k * (bm + 48 * n) + 16 * n
        a2 = (k - 1) * (bm + 40 * n) + 24 * n
        print(f"m={m} n={n} p={inst.p} gen {tg:.1f}s build {tb:.1f}s k={k}: pass1 {tm['pass_one_ms']:.3f} ms "
              f"({a1/tm['pass_one_ms']/1e6:.0f} GB/s, {tm['pass_one_ms']/k*1e3:.1f} us/step)  pass2 {tm['pass_two_ms']:.3f} ms "
              f"({a2/tm['pass_two_ms']/1e6:.0f} GB/s, {tm['pass_two_ms']/max(k-1,1)*1e3:.1f} us/step)  "
              f"total frac {(a1+a2)/((tm['pass_one_ms']+tm['pass_two_ms'])*1e6)/6546.6:.3f}", flush=True)
    del op
