#!/usr/bin/env python
"""Scalability sweep (mirrors the reference's `scalability` binary, src/bin/scalability.rs: arcs sweep at fixed k,
two-pass variant, one CSV row per instance) on 1..8 B200s.

    python scripts/scalability.py --arcs 5000000 20000000 50000000 --k 500
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/scalability.py --arcs 50000000

Columns: arcs,nodes,n,k,gpus,variant,time_s,pass1_ms,pass2_ms,algorithmic_gb,gbs,frac_of_measured_hbm_peak,residual
(time = max over ranks, device-timed; the roofline denominator is gpus x MEASURED_PEAKS.json hbm_gbs)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arcs", type=int, nargs="+", default=[5_000_000, 20_000_000, 50_000_000])
    ap.add_argument("--k", type=int, default=500)
    ap.add_argument("--rho", type=int, default=3)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import two_pass_lanczos_b200 as tpl
    from two_pass_lanczos_b200 import datagen, sharding

    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        peak = 6650.0
    if rank == 0:
        print("arcs,nodes,n,k,gpus,variant,time_s,pass1_ms,pass2_ms,algorithmic_gb,gbs,frac_of_measured_hbm_peak,residual",
              flush=True)
    for m in args.arcs:
        import time

        inst = datagen.gen_kkt(m, args.rho, 1, "aa")
        t_build = time.time()
        if world > 1:
            ident = sharding.broadcast_unique_id(dist, rank)
            op = sharding.sharded_linop(inst.m, inst.p, inst.tail, inst.head, inst.d, rank, world, ident, device=local,
                                       dist=None if os.environ.get("TPL_SHARDED_NCCL") else dist)
        else:
            op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d, device=local)
        op.set_stream(torch.cuda.current_stream().cuda_stream)
        if rank == 0:
            print(f"# {m} arcs: operator built in {time.time() - t_build:.2f} s ({op.kernel_shape()})", file=sys.stderr, flush=True)
        nloc = op.nrows()
        mloc = nloc - inst.p
        b = op.apply(torch.full((nloc,), 1.0 / np.sqrt(inst.n), dtype=torch.float64, device=dev))
        best = None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x = tpl.lanczos_two_pass(op, b, args.k, "inv")
            e1.record()
            e1.synchronize()
            tm = op.last_timing()
            t = torch.tensor([e0.elapsed_time(e1), tm["pass_one_ms"], tm["pass_two_ms"]], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if best is None or float(t[0]) < best[0]:
                best = [float(v) for v in t]
        r = op.apply(x) - b
        sq = torch.stack([(r[:mloc] ** 2).sum(), (b[:mloc] ** 2).sum()])
        if world > 1:
            dist.all_reduce(sq)
        res = float(torch.sqrt((sq[0] + (r[mloc:] ** 2).sum()) / (sq[1] + (b[mloc:] ** 2).sum())))
        bm = 24 * inst.m + 4 * inst.p
        k = args.k
        total = k * (bm + 48 * inst.n) + (k - 1) * (bm + 40 * inst.n) + 40 * inst.n
        if rank == 0:
            gbs = total / (best[0] * 1e-3) / 1e9
            print(f"{inst.m},{inst.p},{inst.n},{k},{world},two-pass,{best[0]*1e-3:.6f},{best[1]:.3f},{best[2]:.3f},"
                  f"{total/1e9:.3f},{gbs:.1f},{gbs/(peak*world):.4f},{res:.3e}", flush=True)
        del op, b, x, r
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
