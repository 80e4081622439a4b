"""Operators around the size where the chunk-resident kernels stop fitting (shared-memory margins): construction and a short
solve must succeed at every size, whatever shape is chosen.  usage: edge_sweep.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import datagen  # noqa: E402

for m in list(range(600_000, 720_001, 8_000)):
    inst = datagen.gen_kkt(m, 3, 1, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    shapes = []
    for mode in (0, 4, 2, 5, 1):
        try:
            op.set_mode(mode)
            x = tpl.lanczos_two_pass(op, b, 6, "inv")
            shapes.append(f"{mode}:{op.kernel_shape()}:{np.linalg.norm(x):.6e}")
        except Exception as e:  # noqa: BLE001
            shapes.append(f"{mode}:ERR {str(e)[:60]}")
    print(m, inst.p, " ".join(shapes), flush=True)
    op.close()
