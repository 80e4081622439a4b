#!/usr/bin/env python
"""Memory / time trade-off sweep (mirrors the reference's `tradeoff` binary, src/bin/tradeoff.rs, and its CSV schema
`variant,k,time_s,rss_kb` -- results/tradeoff_arcs500k_rho3.csv): f(A)b = A^{-1} b on a synthetic netgen-shaped KKT
instance, one-pass (`standard`, keeps V_k in HBM) against two-pass, k = k_start .. k_end.

    python scripts/tradeoff.py [--arcs 500000] [--k-start 50] [--k-end 1000] [--k-step 50] [--output out.csv]

time_s is the wall time of the whole solve through the public API with host buffers; rss_kb is replaced by the bytes of
HBM the handle holds (operator + workspace + basis), in KiB, which is what the one-pass variant trades for its speed here."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import two_pass_lanczos_b200 as tpl  # noqa: E402
from two_pass_lanczos_b200 import datagen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arcs", type=int, default=500_000)
    ap.add_argument("--rho", type=int, default=3)
    ap.add_argument("--k-start", type=int, default=50)
    ap.add_argument("--k-end", type=int, default=1000)
    ap.add_argument("--k-step", type=int, default=50)
    ap.add_argument("--output", default="")
    ap.add_argument("--sweep", action="store_true",
                    help="also time the batched sweeps (lanczos_sweep / lanczos_two_pass_sweep: one basis generation / one pass 1 for all k)")
    args = ap.parse_args()
    inst = datagen.gen_kkt(args.arcs, args.rho, 1, "aa")
    rows = ["variant,k,time_s,rss_kb,kernel_shape,pass1_ms,pass2_or_gemv_ms,rel_diff_vs_two_pass"]
    for k in range(args.k_start, args.k_end + 1, args.k_step):
        x2 = None
        for variant in ("two-pass", "standard"):
            op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)  # fresh handle: no basis left over
            b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))                 # src/bin/tradeoff.rs:234-236
            solve = (lambda: tpl.lanczos_two_pass(op, b, k, "inv")) if variant == "two-pass" else (lambda: tpl.lanczos(op, b, k, "inv"))
            solve()
            t = time.perf_counter()
            x = solve()
            dt = time.perf_counter() - t
            tm = op.last_timing()
            second = tm["pass_two_ms"] if variant == "two-pass" else tm["gemv_ms"]
            diff = 0.0 if x2 is None else float(np.linalg.norm(x - x2) / np.linalg.norm(x2))
            x2 = x if x2 is None else x2
            rows.append(f"{variant},{k},{dt:.9f},{op.device_bytes() // 1024},{op.kernel_shape()},{tm['pass_one_ms']:.3f},"
                        f"{second:.3f},{diff:.2e}")
            print(rows[-1], flush=True)
            op.close()
    if args.sweep:
        ks = list(range(args.k_start, args.k_end + 1, args.k_step))
        per_k = {v: sum(float(r.split(",")[2]) for r in rows[1:] if r.startswith(v + ",")) for v in ("two-pass", "standard")}
        for variant, fn in (("two-pass-sweep", tpl.lanczos_two_pass_sweep), ("standard-sweep", tpl.lanczos_sweep)):
            op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
            b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
            fn(op, b, ks, "inv")
            t = time.perf_counter()
            X = fn(op, b, ks, "inv")
            dt = time.perf_counter() - t
            ref = per_k["two-pass" if variant.startswith("two") else "standard"]
            rows.append(f"{variant},{ks[0]}..{ks[-1]},{dt:.9f},{op.device_bytes() // 1024},{op.kernel_shape()},,,"
                        f"{len(ks)} solutions in one call; the per-k solves above take {ref:.6f} s in total ({ref / dt:.1f}x)")
            print(rows[-1], flush=True)
            del X
            op.close()
    if args.output:
        with open(args.output, "w") as f:
            f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
