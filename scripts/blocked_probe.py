"""Bring-up / timing probe of the blocked streaming kernels (not a test):

    python scripts/blocked_probe.py [--sizes 1000000,5000000,20000000] [--k 100] [--modes 5,2] [--check]

--check: parity of mode 5 (blocked) against mode 2 (tiled) and the oracle on a 60k-arc instance (alpha/beta, regenerated basis,
x).  Timing: per size and mode one warm-up solve and two timed two-pass solves (device events of the library); prints ms per
Lanczos step of each pass and the fraction of the measured HBM peak (algorithmic bytes)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import algorithms as alg, datagen  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


def check():
    import helpers
    from oracle import np_oracle as npo, oracle as orc

    for m, flavour in ((60_000, "wc"), (300_000, "aa")):
        inst = datagen.gen_kkt(m, 3, 11, flavour)
        oop = helpers.oracle_op(inst)
        b = helpers.seeded_b(inst.n)
        k = 40
        ops = {}
        for mode in (5, 2):
            op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
            op.set_mode(mode)
            ops[mode] = op
        print(m, flavour, "shapes", {mo: o.kernel_shape() for mo, o in ops.items()}, flush=True)
        d_ref = orc.lanczos_pass_one(oop, b, k)
        d5 = alg.lanczos_pass_one(ops[5], b, k)
        d2 = alg.lanczos_pass_one(ops[2], b, k)
        sa, sb = np.abs(d_ref.alphas).max(), np.abs(d_ref.betas).max()
        J = 12
        print("  steps", d5.steps_taken, d2.steps_taken, d_ref.steps_taken, "bnorm rel", abs(d5.b_norm - d_ref.b_norm) / d_ref.b_norm)
        print("  alpha rel (first 12) blocked", np.max(np.abs(d5.alphas[:J] - d_ref.alphas[:J])) / sa, "tiled",
              np.max(np.abs(d2.alphas[:J] - d_ref.alphas[:J])) / sa)
        print("  beta  rel (first 12) blocked", np.max(np.abs(d5.betas[:J] - d_ref.betas[:J])) / sb, "tiled",
              np.max(np.abs(d2.betas[:J] - d_ref.betas[:J])) / sb, flush=True)
        out = alg.lanczos_standard(ops[5], b, k)
        y = 0.1 * (np.arange(k) + 1)
        p2 = alg.lanczos_pass_two_with_basis(ops[5], b, d5, y)
        print("  one-pass coefficients == pass-one:", np.array_equal(out.decomposition.alphas, d5.alphas),
              " regenerated basis bit-identical:", np.array_equal(out.v_k, p2.v_k), " x vs V y", helpers.rel(p2.x_k, out.v_k @ y))
        x_plain = alg.lanczos_pass_two(ops[5], b, d5, y)
        print("  pass two without basis == with basis:", np.array_equal(x_plain, p2.x_k))
        if flavour == "wc":
            bn = b / np.linalg.norm(b)
            x5 = tpl.lanczos_two_pass(ops[5], bn, 30, "exp")
            print("  exp k=30 vs oracle", helpers.rel(x5, orc.lanczos_two_pass(oop, bn, 30, npo.exp_tk_solver)), flush=True)
        bc = ops[5].apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
        x5 = tpl.lanczos_two_pass(ops[5], bc, 150, "inv")
        x2 = tpl.lanczos_two_pass(ops[2], bc, 150, "inv")
        print("  inv k=150 blocked vs tiled", helpers.rel(x5, x2), "residual", np.linalg.norm(ops[5].apply(x5) - bc) / np.linalg.norm(bc),
              "deterministic", np.array_equal(tpl.lanczos_two_pass(ops[5], bc, 150, "inv"), x5), flush=True)
        for o in ops.values():
            o.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,5000000,20000000")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--modes", default="5,2")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--shuffle", action="store_true", help="arcs in random order instead of grouped by tail")
    a = ap.parse_args()
    if a.check:
        check()
    pk = peak()
    for m in [int(x) for x in a.sizes.split(",") if x]:
        inst = datagen.gen_kkt(m, 3, 1, "aa")
        tail, head, d = inst.tail, inst.head, inst.d
        if a.shuffle:
            perm = np.random.default_rng(0).permutation(m)
            tail, head, d = tail[perm], head[perm], d[perm]
        t0 = time.time()
        op = tpl.LinOp.from_kkt(inst.m, inst.p, tail, head, d)
        build_s = time.time() - t0
        b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
        n, k = inst.n, a.k
        bm = 24 * inst.m + 4 * inst.p
        a1, a2 = k * (bm + 48 * n) + 16 * n, (k - 1) * (bm + 40 * n) + 24 * n
        for mode in [int(x) for x in a.modes.split(",")]:
            op.set_mode(mode)
            best = None
            for rep in range(3):
                x = tpl.lanczos_two_pass(op, b, k, "inv")
                tm = op.last_timing()
                if rep and (best is None or tm["pass_one_ms"] + tm["pass_two_ms"] < best[0] + best[1]):
                    best = (tm["pass_one_ms"], tm["pass_two_ms"])
            res = np.linalg.norm(op.apply(x) - b) / np.linalg.norm(b)
            print(f"m={m} k={k} mode={mode} {op.kernel_shape():8s} build {build_s:.2f}s  pass1 {best[0]:9.3f} ms ({1e3 * best[0] / k:7.2f} us/step, "
                  f"{a1 / best[0] / 1e6 / pk:.3f} of peak)  pass2 {best[1]:9.3f} ms ({1e3 * best[1] / max(k - 1, 1):7.2f} us/step, "
                  f"{a2 / best[1] / 1e6 / pk:.3f})  total {(a1 + a2) / (best[0] + best[1]) / 1e6 / pk:.3f}  residual {res:.2e}", flush=True)
        op.close()


if __name__ == "__main__":
    main()
