#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native two-pass Lanczos engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--arcs M] [--k K_LANCZOS]

Metric (BASELINE.json): wall time of `lanczos_two_pass` f(A)b, f = inv, on a synthetic netgen-shaped KKT
instance with 500k arcs, rho = 3 (n = 501 155), k = 500, b = A*(1/sqrt(n))*1 (src/bin/tradeoff.rs:234-258).
One "step" = one complete two-pass solve (pass 1, host f(T_k) e1, pass 2).

Prints ONE JSON line (rank 0).  `value` = device-timed ms per solve with b and x resident in HBM; `e2e` = the
same solve through the public host API with HOST buffers (H2D of b and D2H of x inside the timed region);
`roofline` = algorithmic bytes of the dominant kernel (pass 1) / its measured duration against the measured HBM
peak; `cpu_baseline` = the CPU oracle (a faer-faithful single-thread port of the reference) timed on this box.
`--impl reference` times that CPU port alone (the Rust/faer reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "two-pass f(A)b wall time (500k arcs, rho=3, k=500)"
REF_PUBLISHED_S = 5.275  # results/scalability_k500_rho3.csv:21 (Xeon Gold 5318Y, 1 core) -- other hardware


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arcs", type=int, default=500_000)
    ap.add_argument("--rho", type=int, default=3)
    ap.add_argument("--k", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--large-arcs", type=int, default=50_000_000,
                    help="arcs of the large streaming instance timed after the headline workload (0 = skip)")
    return ap.parse_args()


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("pass1_kernel_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            return None
    return None


def recorded_large_traffic(shape, arcs):
    """dram bytes per Lanczos step of the streaming kernels on the large instance, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if shape == "blocked" and arcs == 50_000_000 and os.path.exists(path):
        try:
            return json.load(open(path)).get("blocked_50M_arcs")
        except Exception:  # noqa: BLE001
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_instance(args):
    from two_pass_lanczos_b200 import datagen

    inst = datagen.gen_kkt(args.arcs, args.rho, args.seed, "aa")
    return inst


def config_dict(args, inst):
    """`config` of the JSON line: identical in both arms (same workload, same N)."""
    world = args.gpus
    return {"workload": workload_string(args, inst),
            "l2": "GPU arm: L2 flushed between solves (256 MiB write); CPU arm: 28 MB working set, cold first run discarded",
            "parallelism": "1 GPU" if world == 1 else
            f"{world} GPUs behind the arc-partitioned rank-local API (rank r owns an arc block + a replica of the p={inst.p} node "
            f"entries); the library shards the solve when that pays (per step a reduction of the node sums and two scalar "
            f"all-reduces, fused into the kernels) and solves replicated when the whole job fits one GPU's on-chip kernels "
            f"(`arm.kernel_shape`); the CPU arm runs on rank 0's host cores"}


def workload_string(args, inst):
    """ONE description of the workload, shared verbatim by both arms (the driver compares the two lines' `config.workload`)."""
    return (f"netgen-shaped KKT {args.arcs} arcs rho={args.rho} n={inst.n} (seed {args.seed}, qfcgen 'aa' costs), "
            f"lanczos_two_pass f=inv k={args.k}, b=A*(1/sqrt(n))")


def algorithmic_bytes(n, bmat, k):
    """pass-1 step = B_mat + 48 n, pass-2 step = B_mat + 40 n, init = 40 n."""
    pass1 = k * (bmat + 48 * n) + 16 * n
    pass2 = (k - 1) * (bmat + 40 * n) + 24 * n
    return pass1, pass2


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_model():
    """CPU model and core count of the box (stated next to every CPU number)."""
    name = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    name = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return f"{name} ({os.cpu_count()} logical cores, 1 used)"


def cpu_two_pass_seconds(inst, k, repeats=1, keep=None):
    """Times the CPU oracle (oracle/lanczos_oracle.cpp: CSC scatter matvec, 8-byte indices, unfused sweeps, one
    thread -- the reference runs faer with Par::Seq) on the same instance and right-hand side.  `keep` (a dict) receives the
    oracle's x and b: the GPU arm compares its own x with it (`parity` in the JSON line), outside every timed region."""
    from oracle import np_oracle as npo
    from oracle import oracle as orc

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers

    oop = helpers.oracle_op(inst)
    b = oop.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        x = orc.lanczos_two_pass(oop, b, k, npo.inv_tk_solver)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        if keep is not None:
            keep["x"], keep["b"] = x, b
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    inst = build_instance(args)
    total = args.steps + args.warmup
    # bounded sample: the CPU cost is linear in k (2k-1 matvecs); keep the whole run within ~150 s
    t_probe = cpu_two_pass_seconds(inst, 10)
    per_k = t_probe / 10.0
    k_s = args.k
    if per_k * args.k * total > 150.0:
        k_s = max(10, int(150.0 / (per_k * total)))
    times = []
    for i in range(total):
        dt = cpu_two_pass_seconds(inst, k_s)
        if i >= args.warmup:
            times.append(dt * (args.k / k_s))
    ms = 1e3 * sum(times) / len(times)
    sample = (f"full workload (k={args.k})" if k_s == args.k else
              f"k={k_s} of {args.k} Lanczos steps per solve, time scaled by {args.k}/{k_s} (cost is linear in k)")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, inst),
        "arm": {"format": "csc (faer SparseColMat, 8-byte indices)",
                "note": "reference = CPU oracle port (Rust/faer reference cannot be built here: no cargo/rustc, faer "
                        "un-vendored); single thread because the reference runs every faer call with Par::Seq"},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": sample, "cpu": cpu_model()},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "published_reference_ms_other_hw": REF_PUBLISHED_S * 1e3,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def csr_path_leg(args, inst, b_host, x_incidence):
    """The headline workload through the GENERIC CSR operator (north star "Matvec": generic CSR kernel next to the specialised
    incidence kernel; what `&kkt.a` is to the reference, src/algorithms/mod.rs:177): SELL-32 + long-row kernels of
    tpl_csr.cuh on the host CSC of the KKT matrix.  Device-timed, best of two solves; byte model B_csr = 12 nnz + 4 (n + 1)."""
    import two_pass_lanczos_b200 as tpl
    from two_pass_lanczos_b200 import datagen

    cp, ri, va = datagen.kkt_csc(inst)
    op = tpl.LinOp.from_csc(inst.n, cp, ri, va)
    k, n = args.k, inst.n
    best, x = None, None
    for _ in range(3):
        x = tpl.lanczos_two_pass(op, b_host, k, "inv")
        tm = op.last_timing()
        if best is None or tm["pass_one_ms"] + tm["pass_two_ms"] < sum(best):
            best = (tm["pass_one_ms"], tm["pass_two_ms"])
    a1, a2 = algorithmic_bytes(n, op.matrix_bytes(), k)
    peak, _ = measured_peak_gbs()
    out = {"workload": "the headline instance as a generic CSR operator (tpl_op_from_csc)", "nnz": int(len(va)),
           "kernel_shape": op.kernel_shape(), "time_ms": sum(best), "pass1_ms": best[0], "pass2_ms": best[1],
           "matrix_bytes_model": int(op.matrix_bytes()), "frac_of_hbm_peak": (a1 + a2) / (sum(best) * 1e-3) / 1e9 / peak,
           "x_rel_vs_incidence_path": float(np.linalg.norm(x - x_incidence) / np.linalg.norm(x_incidence)),
           "note": "2.5 M non-zeros: the matrix and the vectors sit in L2; fixed per-step costs (two grid barriers) dominate"}
    op.close()
    return out


def large_instance_leg(args, rank, world, local_rank, dist, dev, shuffled=False):
    """BASELINE.json config 4 / north_star: the large synthetic instance (default 50M arcs, rho = 3, k = 500) on the same
    N GPUs, after the headline workload -- a reported extra (`large_instance` in the JSON line), not the headline value.
    One GPU: tiled streaming kernels; N > 1: arc-partitioned, the same kernels spanning all ranks."""
    import torch

    import two_pass_lanczos_b200 as tpl
    from two_pass_lanczos_b200 import datagen, sharding

    inst = datagen.gen_kkt(args.large_arcs, args.rho, args.seed, "aa")
    if shuffled:  # the same graph with its arcs in random order (netgen emits them grouped by tail)
        perm = np.random.default_rng(args.seed).permutation(inst.m)
        inst.tail, inst.head, inst.d = inst.tail[perm], inst.head[perm], inst.d[perm]
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    if world > 1:
        ident = sharding.broadcast_unique_id(dist, rank)
        op = sharding.sharded_linop(inst.m, inst.p, inst.tail, inst.head, inst.d, rank, world, ident, device=local_rank,
                                   dist=None if os.environ.get("TPL_SHARDED_NCCL") else dist)
    else:
        op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d, device=local_rank)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    stream = torch.cuda.current_stream()
    op.set_stream(stream.cuda_stream)
    nloc = op.nrows()
    mloc = nloc - inst.p
    b = op.apply(torch.full((nloc,), 1.0 / np.sqrt(inst.n), dtype=torch.float64, device=dev))
    best = None
    for _ in range(2):  # the first solve warms up; vectors (2.4 GB) are far larger than L2
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        x = tpl.lanczos_two_pass(op, b, args.k, "inv")
        e1.record(stream)
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
    a1, a2 = algorithmic_bytes(inst.n, 24 * inst.m + 4 * inst.p, args.k)
    peak, _ = measured_peak_gbs()
    gbs = (a1 + a2) / (best[0] * 1e-3) / 1e9
    out = {"workload": f"netgen-shaped KKT {inst.m} arcs rho={args.rho} n={inst.n}, lanczos_two_pass f=inv k={args.k}",
           "n_gpus": world, "time_s": best[0] * 1e-3, "pass1_ms": best[1], "pass2_ms": best[2],
           "algorithmic_gb": (a1 + a2) / 1e9, "gbs": gbs, "frac_of_hbm_peak": gbs / (peak * world),
           "pass1_frac": a1 / (best[1] * 1e-3) / 1e9 / (peak * world), "pass2_frac": a2 / (best[2] * 1e-3) / 1e9 / (peak * world),
           "arc_order": "random permutation" if shuffled else "grouped by tail (netgen order)",
           "operator_build_s": round(build_s, 3),
           "kernel_shape": op.kernel_shape(), "residual": res, "traffic": recorded_large_traffic(op.kernel_shape(), inst.m),
           "timing": "CUDA events on the launching stream, max over ranks, best of 2 solves"}
    op.close()
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import two_pass_lanczos_b200 as tpl

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    inst = build_instance(args)
    k = args.k
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    if world > 1:
        # arc-partitioned operator: rank r owns a contiguous arc block + a replica of the node entries; the
        # per-step all-reduces run inside the library over NCCL (same NVLink fabric torch.distributed uses)
        from two_pass_lanczos_b200 import sharding

        ident = sharding.broadcast_unique_id(dist, rank)
        op = sharding.sharded_linop(inst.m, inst.p, inst.tail, inst.head, inst.d, rank, world, ident, device=local_rank,
                                   dist=None if os.environ.get("TPL_SHARDED_NCCL") else dist)
        x_true = torch.full((op.nrows(),), 1.0 / np.sqrt(inst.n), dtype=torch.float64, device=dev)
    else:
        op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d, device=local_rank)
        x_true = torch.full((inst.n,), 1.0 / np.sqrt(inst.n), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    n = op.nrows()          # rank-local vector length (= inst.n on one GPU)
    m_loc = n - inst.p
    op.set_stream(stream.cuda_stream)
    b_dev = op.apply(x_true)
    b_host = torch.empty(n, dtype=torch.float64).pin_memory()
    b_host.copy_(b_dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve_dev():
        return tpl.lanczos_two_pass(op, b_dev, k, "inv")

    def solve_host():
        return tpl.lanczos_two_pass(op, b_host, k, "inv")

    # nvidia-smi needs ~0.1 s before its first sample: it is started before the warm-up so that it is already sampling
    # (every 50 ms) when the timed regions begin; the clocks reported are the median over warm-up + timed regions
    sampler = ClockSampler(local_rank)
    sampler.start()
    warm = []
    for _ in range(args.warmup):
        solve_dev()
        warm.append(solve_host())  # results kept until the warm-up ends: the pinned-memory cache then holds enough blocks
    del warm
    # ---- value: device-resident inputs, CUDA events on the launching stream, L2 flushed between solves
    launches0 = op.kernel_launches()
    barrier()
    dev_ms, p1_ms, p2_ms = [], [], []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        x_dev = solve_dev()
        e1.record(stream)
        e1.synchronize()
        dev_ms.append(e0.elapsed_time(e1))
        tm = op.last_timing()
        p1_ms.append(tm["pass_one_ms"])
        p2_ms.append(tm["pass_two_ms"])
    barrier()
    launches = op.kernel_launches() - launches0
    # ---- e2e: host buffers through the public API (H2D b, D2H x inside the timed region)
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        x_host = solve_host()
        e2e_ms.append((time.perf_counter() - t) * 1e3)
    barrier()
    clocks = sampler.stop()

    t_dev = torch.tensor([sum(dev_ms) / len(dev_ms), sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms, e2e = float(t_dev[0]), float(t_dev[1])

    # residual ||A x - b|| / ||b|| of the last solve (arc parts summed over ranks, replicated node part counted once)
    r_loc = op.apply(x_dev) - b_dev
    sq = torch.stack([(r_loc[:m_loc] ** 2).sum(), (b_dev[:m_loc] ** 2).sum()])
    if world > 1:
        dist.all_reduce(sq)
    res = float(torch.sqrt((sq[0] + (r_loc[m_loc:] ** 2).sum()) / (sq[1] + (b_dev[m_loc:] ** 2).sum())))
    shape = op.kernel_shape()
    # ---- parity: the whole x of the last device-resident solve on rank 0 (arc slices in rank order + the node replica)
    if world > 1:
        from two_pass_lanczos_b200 import sharding

        lens = [sharding.arc_range(inst.m, r, world) for r in range(world)]
        pad = max(hi - lo for lo, hi in lens) + inst.p
        mine = torch.zeros(pad, dtype=torch.float64, device=dev)
        mine[:n] = x_dev
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        x_full = np.concatenate([parts[r][: hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(lens)]
                                + [parts[0][lens[0][1] - lens[0][0]: lens[0][1] - lens[0][0] + inst.p].cpu().numpy()])
        node_replicas_equal = all(
            np.array_equal(parts[r][hi - lo: hi - lo + inst.p].cpu().numpy(), x_full[inst.m:]) for r, (lo, hi) in enumerate(lens))
    else:
        x_full = x_dev.cpu().numpy()
        node_replicas_equal = True
    large = None
    if args.large_arcs > 0:
        try:
            large = large_instance_leg(args, rank, world, local_rank, dist, dev)
            if world == 1:  # the same instance with its arcs in random order: the kernels must not depend on netgen's grouping
                large["shuffled_arcs"] = large_instance_leg(args, rank, world, local_rank, dist, dev, shuffled=True)
        except Exception as e:  # noqa: BLE001 - the extra leg must never take the headline line down
            large = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        assert np.isfinite(res) and res < 1e-6, f"solve did not converge: residual {res}"
        assert np.array_equal(np.asarray(x_host), x_dev.cpu().numpy()), "host and device paths disagree"
        bmat = op.matrix_bytes()
        a1, a2 = algorithmic_bytes(inst.n, bmat * world, k)   # whole job (all ranks)
        peak, peak_src = measured_peak_gbs()
        p1 = sum(p1_ms) / len(p1_ms)
        p2 = sum(p2_ms) / len(p2_ms)
        achieved = a1 / (p1 * 1e-3) / 1e9
        kernel_name = {
            "cells": "pass1_cell_kernel<false> (2-D cell partition, arcs in registers; one persistent launch per pass)",
            "chunks": "pass1_resident_kernel<false> (contiguous chunks in shared memory; one persistent launch per pass)",
            "tiled": "pass1_tiled_kernel<false> (streaming, tiled node sums; one persistent launch per pass)",
            "gather": "pass1_kernel<IncidenceOp,false> (streaming, gathered node rows; one persistent launch per pass)",
            "blocked": "pass1_blocked_kernel<false> (streaming, 2-D node-block partition, cell-order vectors, bulk-copy input ring; "
                       "one persistent launch per pass)",
            "replicated": "pass1_cell_kernel<false> on every rank (the whole operator fits the on-chip cell kernels: the ranks "
                          "all-reduce their slices of b once and solve redundantly, no per-step communication)",
            "sharded": "shard_phase_a_kernel + shard_phase_b_kernel (2 launches + 2 NCCL all-reduces per step)",
            "sharded-blocked": "pass1_blocked_kernel<false> spanning all ranks (destination-indexed node-sum exchange, node-value "
                               "all-gather and alpha / beta all-reduce as peer-memory stores over NVLink inside one persistent "
                               "launch per pass)",
            "sharded-fused": "pass1_tiled_kernel<false> spanning all ranks (node-sum reduce-scatter, node-value all-gather and "
                             "alpha / beta all-reduce as peer-memory stores over NVLink inside one persistent launch per pass)",
        }.get(shape, shape)
        on_chip = shape in ("cells", "chunks", "replicated")
        line = {
            "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, inst),
            "arm": {"format": "incidence", "kernel_shape": shape, "operator_build_s": round(build_s, 4),
                    "operator_build_note": "operator tables (host builders below 2^20 arcs, on the device above) + H2D of the operator, outside the timed region as in the reference's "
                                           "protocol (src/bin/tradeoff.rs:265-288 times the solver call only)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "ms", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n + 16 * k,
                    "per_step_ms": [round(t, 3) for t in e2e_ms]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak * world,
                         "unit": "GB/s", "frac": achieved / (peak * world), "traffic": recorded_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": a1, "kernel_ms": p1,
                         "pass2_kernel_ms": p2, "pass2_achieved": a2 / (p2 * 1e-3) / 1e9,
                         "whole_solve_frac": (a1 + a2) / (ms * 1e-3) / 1e9 / (peak * world),
                         "note": ("operator and vectors stay in registers / shared memory for the whole pass: `achieved` is "
                                  "algorithmic bytes / time, an EFFECTIVE bandwidth that HBM never carries (ncu dram "
                                  "traffic per launch is in `traffic`)") if on_chip else
                                 "streams the operator and the vectors from HBM every step"},
            "residual": res,
            "published_reference_ms_other_hw": REF_PUBLISHED_S * 1e3,
        }
        if large is not None:
            line["large_instance"] = large
        if world == 1:
            try:
                line["csr_path"] = csr_path_leg(args, inst, b_host.numpy().copy(), x_full)
            except Exception as e:  # noqa: BLE001 - the extra leg must never take the headline line down
                line["csr_path"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline:
            # the CPU oracle solves the same instance and b once (outside every timed region): its time is the cpu_baseline
            # (reported at N = 1 only), its x is what `parity` compares the GPU result with -- at every N
            kept = {}
            t_cpu = cpu_two_pass_seconds(inst, k, keep=kept)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers

            x_cpu = kept["x"]
            line["parity"] = {
                "x_rel_vs_cpu": helpers.rel(x_full, x_cpu),
                "x_rel_vs_cpu_nullspace_projected": helpers.rel(helpers.project_out_null(x_full, inst.m, inst.p),
                                                                 helpers.project_out_null(x_cpu, inst.m, inst.p)),
                "tolerance": 1e-10, "node_replicas_bit_equal": bool(node_replicas_equal),
                "what": f"||x_gpu - x_cpu|| / ||x_cpu||, x_gpu = the last device-timed solve assembled from {world} rank(s), "
                        f"x_cpu = oracle/lanczos_oracle.cpp lanczos_two_pass on the same instance and b (k={k}, f=inv)"}
            assert line["parity"]["x_rel_vs_cpu"] <= 1e-10, f"parity gate failed: {line['parity']}"
            if world == 1:
                line["cpu_baseline"] = {"value": t_cpu * 1e3, "unit": "ms", "cores": 1, "kind": "port", "cpu": cpu_model(),
                                        "sample": f"full workload (k={k}), one cold run, single thread (Par::Seq)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from two_pass_lanczos_b200 import build as tpl_build

    if local_rank == 0:
        tpl_build.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        time.sleep(0.5 if local_rank else 0.0)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
