"""Quick GPU bring-up script (not a test): prints parity and timing diagnostics."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen, data_loader  # noqa: E402
from oracle import oracle as orc, np_oracle as npo  # noqa: E402
import helpers  # noqa: E402


def main():
    big = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
    A = np.array([[2, -1, 0, 0], [-1, 2, -1, 0], [0, -1, 2, -1], [0, 0, -1, 2.0]])
    op = tpl.LinOp.from_dense(A)
    print("apply", op.apply(np.arange(1, 5.0)), A @ np.arange(1, 5.0))
    d = alg.lanczos_pass_one(op, [1, 0, 0, 0], 2)
    print("KAT", d)
    op2 = tpl.LinOp.from_dense(np.diag([2.0, 3.0]))
    out = alg.lanczos_standard(op2, [1, 0], 2)
    print("breakdown", out.decomposition, out.v_k.shape)
    try:
        alg.lanczos_standard(tpl.LinOp.from_dense(np.eye(2)), [0, 0], 2)
    except tpl.LanczosError as e:
        print("zero b:", e.kind, e)
    b = np.arange(1, 5.0)
    x1 = tpl.lanczos(op, b, 3, npo.inv_tk_solver)
    x2 = tpl.lanczos_two_pass(op, b, 3, "inv")
    print("doctest", np.linalg.norm(x1 - x2), x1, np.linalg.solve(A, b))

    # golden 1000-arc netgen instance, line-per-value qfc written from the generator-free path
    for m, k in ((1000, 30), (50000, 100)):
        inst = datagen.gen_kkt(m, 3, 7, "wc")
        oop = helpers.oracle_op(inst)
        bb = helpers.seeded_b(inst.n)
        od = orc.lanczos_pass_one(oop, bb, k)
        for fmt in ("incidence", "csr"):
            if fmt == "incidence":
                gop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
            else:
                cp, ri, va = datagen.kkt_csc(inst)
                gop = tpl.LinOp.from_csc(inst.n, cp, ri, va)
            y = gop.apply(bb)
            print(m, fmt, "apply rel", helpers.rel(y, oop.apply(bb)))
            gd = alg.lanczos_pass_one(gop, bb, k)
            print(m, fmt, "steps", gd.steps_taken, od.steps_taken, "bnorm", gd.b_norm - od.b_norm,
                  "alpha rel", np.max(np.abs(gd.alphas - od.alphas) / np.abs(od.alphas).max()),
                  "beta rel", np.max(np.abs(gd.betas - od.betas) / np.abs(od.betas).max()))
            so = alg.lanczos_standard(gop, bb, k)
            yk = 0.1 * (np.arange(gd.steps_taken) + 1)
            p2 = alg.lanczos_pass_two_with_basis(gop, bb, gd, yk)
            print(m, fmt, "drift", np.abs(so.v_k - p2.v_k).max(), "alpha one-pass vs pass-one",
                  np.abs(so.decomposition.alphas - gd.alphas).max(), "ortho",
                  np.linalg.norm(np.eye(k) - so.v_k.T @ so.v_k))
            xo = orc.lanczos_pass_two(oop, bb, od, yk)
            print(m, fmt, "x rel vs oracle", helpers.rel(p2.x_k, xo))
            xg1 = tpl.lanczos(gop, bb, k, "exp")
            xg2 = tpl.lanczos_two_pass(gop, bb, k, "exp")
            xoe = orc.lanczos_two_pass(oop, bb, k, npo.exp_tk_solver)
            print(m, fmt, "exp one-pass vs two-pass", helpers.rel(xg1, xg2), "vs oracle", helpers.rel(xg2, xoe))

    # headline shape
    inst = datagen.gen_kkt(big, 3, 1, "aa")
    gop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    bb = gop.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    for k in (50, 500):
        for rep in range(3):
            t = time.time()
            x = tpl.lanczos_two_pass(gop, bb, k, "inv")
            wall = time.time() - t
            tm = gop.last_timing()
            print(f"m={big} k={k} wall {wall*1e3:.2f} ms  pass1 {tm['pass_one_ms']:.3f} ms pass2 {tm['pass_two_ms']:.3f} ms")
    for mode in (2, 1, 0):
        gop.set_mode(mode)
        xm = tpl.lanczos_two_pass(gop, bb, 500, "inv")
        tm = gop.last_timing()
        print(f"mode {mode}: pass1 {tm['pass_one_ms']:.3f} ms pass2 {tm['pass_two_ms']:.3f} ms  bitwise equal to mode 0: {np.array_equal(xm, x)}")
        dm = alg.lanczos_pass_one(gop, bb, 60)
        if mode == 2:
            d2 = dm
        print("   alpha/beta equal to mode 2:", np.array_equal(dm.alphas, d2.alphas), np.array_equal(dm.betas, d2.betas))
    n = inst.n
    bm = gop.matrix_bytes()
    k = 500
    total = k * (bm + 48 * n) + (k - 1) * (bm + 40 * n) + 40 * n
    tm = gop.last_timing()
    sec = (tm['pass_one_ms'] + tm['pass_two_ms']) * 1e-3
    print("algorithmic GB", total / 1e9, "GB/s", total / sec / 1e9, "frac of 6546.6:", total / sec / 6546.6e9)
    res = gop.apply(x) - bb
    print("residual", np.linalg.norm(res) / np.linalg.norm(bb))
    t = time.time()
    x1 = tpl.lanczos(gop, bb, 500, "inv")
    print("one-pass wall", time.time() - t, gop.last_timing(), "dev", helpers.rel(x1, x))


if __name__ == "__main__":
    main()
