import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import two_pass_lanczos_b200 as tpl
from oracle import np_oracle as npo, oracle as orc
from two_pass_lanczos_b200 import datagen
for seed in (9, 10, 11):
    inst = datagen.gen_kkt(1000, 3, seed, "wc")
    oop = helpers.oracle_op(inst)
    b = oop.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    for k in (150, 200, 250):
        xo = orc.lanczos_two_pass(oop, b, k, npo.inv_tk_solver)
        x_ref = helpers.project_out_null(xo, inst.m, inst.p)
        line = [f"seed {seed} k {k} oracle res {np.linalg.norm(oop.apply(xo)-b)/np.linalg.norm(b):.2e}"]
        for name, mode in (("res", 0), ("tiled", 2), ("gather", 3)):
            gop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d); gop.set_mode(mode)
            xg = tpl.lanczos_two_pass(gop, b, k, npo.inv_tk_solver)
            x = helpers.project_out_null(xg, inst.m, inst.p)
            line.append(f"{name}: dev {helpers.rel(x, x_ref):.2e} res {np.linalg.norm(oop.apply(xg)-b)/np.linalg.norm(b):.2e}")
        print(" | ".join(line), flush=True)
