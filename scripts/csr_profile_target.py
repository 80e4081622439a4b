"""Small fixed CSR workload for ncu: two-pass solves through the generic CSR operator.  usage: csr_profile_target.py [arcs] [k]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import datagen  # noqa: E402

arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
inst = datagen.gen_kkt(arcs, 3, 1, "wc")
cp, ri, va = datagen.kkt_csc(inst)
op = tpl.LinOp.from_csc(inst.n, cp, ri, va)
b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
x = tpl.lanczos_two_pass(op, b, k, "inv")
print(op.kernel_shape(), op.last_timing())
