"""Small, fixed workload for ncu: two-pass solves on a synthetic instance.  usage: profile_target.py [k] [arcs] [reps] [mode]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import datagen  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 100
arcs = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
inst = datagen.gen_kkt(arcs, 3, 1, "aa")
op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
op.set_mode(mode)
print(op.kernel_shape())
for _ in range(reps):
    x = tpl.lanczos_two_pass(op, b, k, "inv")
    print(op.last_timing())
print("residual", np.linalg.norm(op.apply(x) - b) / np.linalg.norm(b))
