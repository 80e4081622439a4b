"""Timing + parity probe of the generic CSR path (north_star "Matvec": generic CSR kernel):

    python scripts/csr_probe.py [--sizes 1000000,5000000] [--k 50]

Builds the KKT matrix as the host CSC of `KKTSystem.a` (what `load_kkt_system` hands back, src/utils/data_loader.rs:251), runs
the two-pass solve through the CSR operator and through the incidence operator, and prints ms per Lanczos step, the fraction of
the measured HBM peak against the CSR byte model (B_csr = 12 nnz + 4 (n + 1)) and the deviation of the two x."""
import argparse
import json
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
    ap.add_argument("--sizes", default="1000000,5000000")
    ap.add_argument("--k", type=int, default=50)
    a = ap.parse_args()
    try:
        pk = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pk = 6650.0
    for m in [int(x) for x in a.sizes.split(",") if x]:
        inst = datagen.gen_kkt(m, 3, 1, "wc")
        cp, ri, va = datagen.kkt_csc(inst)
        t0 = time.time()
        op = tpl.LinOp.from_csc(inst.n, cp, ri, va)
        build_s = time.time() - t0
        inc = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
        b = inc.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
        n, k, nnz = inst.n, a.k, len(va)
        bm = 12 * nnz + 4 * (n + 1)
        assert bm == op.matrix_bytes()
        a1, a2 = k * (bm + 48 * n) + 16 * n, (k - 1) * (bm + 40 * n) + 24 * n
        best = None
        for rep in range(3):
            x = tpl.lanczos_two_pass(op, b, k, "inv")
            tm = op.last_timing()
            if rep and (best is None or tm["pass_one_ms"] + tm["pass_two_ms"] < sum(best)):
                best = (tm["pass_one_ms"], tm["pass_two_ms"])
        xi = tpl.lanczos_two_pass(inc, b, k, "inv")
        dev = float(np.linalg.norm(x - xi) / np.linalg.norm(xi))
        print(f"csr m={m} nnz={nnz} k={k} {op.kernel_shape()} build {build_s:.2f}s  pass1 {1e3 * best[0] / k:8.2f} us/step ({a1 / best[0] / 1e6 / pk:.3f} of peak)  "
              f"pass2 {1e3 * best[1] / max(k - 1, 1):8.2f} us/step ({a2 / best[1] / 1e6 / pk:.3f})  total {(a1 + a2) / sum(best) / 1e6 / pk:.3f}  "
              f"x vs incidence path {dev:.2e}", flush=True)
        op.close()
        inc.close()


if __name__ == "__main__":
    main()
