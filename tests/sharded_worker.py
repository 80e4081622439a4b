"""One rank of the multi-GPU parity run (launched by tests/test_gpu_sharded.py through torch.distributed.run).
Writes its results for rank 0 to compare with the unsharded engine and the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import helpers
    import two_pass_lanczos_b200 as tpl
    from two_pass_lanczos_b200 import algorithms as alg
    from two_pass_lanczos_b200 import datagen, sharding

    out_path, m, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    fused = len(sys.argv) > 4 and sys.argv[4] == "fused"
    replicated = len(sys.argv) > 4 and sys.argv[4] == "replicated"  # else TPL_NO_REPLICATE=1 is set by the launcher
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    inst = datagen.gen_kkt(m, 3, 7, "wc")
    p = inst.p
    ident = sharding.broadcast_unique_id(dist, rank)
    op = sharding.sharded_linop(m, p, inst.tail, inst.head, inst.d, rank, world, ident, device=local,
                               dist=dist if fused else None)
    if replicated and world > 1:
        assert op.kernel_shape() == "replicated"
    else:
        assert op.kernel_shape() in (("sharded-blocked", "sharded-fused") if fused and world > 1 else ("sharded",))
    lo, hi = op.arc_lo, op.arc_hi
    info = op.shard_info()
    assert info == {"rank": rank, "world": world, "local_arcs": hi - lo, "nodes": p}
    xg = np.full(inst.n, 1.0 / np.sqrt(inst.n))
    b_loc = op.apply(sharding.local_vector(xg, m, p, lo, hi))          # b = A * const, rank-local layout
    dec = alg.lanczos_pass_one(op, b_loc, k)
    std = alg.lanczos_standard(op, b_loc, k)
    yk = 0.1 * (np.arange(dec.steps_taken) + 1)
    p2 = alg.lanczos_pass_two_with_basis(op, b_loc, dec, yk)
    x_exp = tpl.lanczos_two_pass(op, b_loc, k, "exp")
    x_one = tpl.lanczos(op, b_loc, k, "exp")
    # device tensors on torch's current stream (the handle, and the inner handle of a replicated operator, adopt it)
    x_dev = tpl.lanczos_two_pass(op, torch.from_numpy(b_loc).cuda(), k, "exp")
    assert x_dev.is_cuda and np.array_equal(x_dev.cpu().numpy(), x_exp)
    res = {"b": b_loc, "alphas": dec.alphas, "betas": dec.betas, "b_norm": dec.b_norm, "steps": dec.steps_taken,
           "drift": float(np.abs(std.v_k - p2.v_k).max()), "std_alphas": std.decomposition.alphas, "x_p2": p2.x_k,
           "x_exp": x_exp, "x_one": x_one, "launches": op.kernel_launches()}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        glob = {}
        for key in ("b", "x_p2", "x_exp", "x_one"):
            glob[key] = sharding.assemble_global([g[key] for g in gathered], m, p)
            glob[key + "_replica_gap"] = max(float(np.abs(g[key][len(g[key]) - p:] - gathered[0][key][len(gathered[0][key]) - p:]).max())
                                             for g in gathered)
        for key in ("alphas", "betas", "std_alphas"):
            glob[key] = gathered[0][key]
            glob[key + "_same"] = all(np.array_equal(g[key], gathered[0][key]) for g in gathered)
        glob["b_norm"] = gathered[0]["b_norm"]
        glob["steps"] = gathered[0]["steps"]
        glob["drift"] = max(g["drift"] for g in gathered)
        glob["launches"] = min(g["launches"] for g in gathered)
        np.savez(out_path, **glob)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
