"""Regenerates the committed golden fixtures.  Run from the repo root:  python tests/golden/make_golden.py

1. netgen1000/*.dmx, *.qfc, *.par were produced ONCE with the reference's own prebuilt tools
   (data/qcnd/pargen 1000 {1,2,3} 1 a a ns -> data/netgen/src/netgen -> data/qcnd/qfcgen); pargen seeds from
   time(NULL), so they are committed as files rather than regenerated.  The `.qfc` files have qfcgen's 3-line
   layout, on which the reference loader returns an EMPTY D block.
2. This script derives `*.lines.qfc` (one value per line: the layout the reference loader really parses) from
   them and writes golden_vectors.npz: oracle alpha/beta/x for every instance (oracle = oracle/lanczos_oracle.cpp).
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import np_oracle as npo  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    out = {}
    for dmx in sorted(glob.glob(os.path.join(HERE, "netgen1000", "*.dmx"))):
        base = dmx[:-4]
        name = os.path.basename(base)
        toks = open(base + ".qfc").read().split("\n")
        m = int(toks[0])
        fixed = toks[1].split()
        quad = toks[2].split()
        assert len(fixed) == m and len(quad) == m
        with open(base + ".lines.qfc", "w") as f:
            f.write(f"{m}\n" + "\n".join(fixed) + "\n" + "\n".join(quad) + "\n")
        # "wc" flavour: D ~ U[1,10] (tex/report.tex:338-342), seeded
        rng = np.random.default_rng(1000 + len(out))
        wc = 1.0 + 9.0 * rng.random(m)
        with open(base + ".wc.qfc", "w") as f:
            f.write(f"{m}\n" + "\n".join(fixed) + "\n" + "\n".join(repr(float(v)) for v in wc) + "\n")
        for flavour, qfc in (("nod", base + ".qfc"), ("aa", base + ".lines.qfc"), ("wc", base + ".wc.qfc")):
            kkt = orc.load_kkt_system(dmx, qfc)
            n = kkt.a.nrows()
            b = np.random.default_rng(42).random(n)
            dec = orc.lanczos_pass_one(kkt.a, b, 30)
            key = f"{name}.{flavour}"
            out[key + ".nnz"] = np.array([kkt.a.nnz(), kkt.num_nodes, kkt.num_arcs])
            out[key + ".b"] = b
            out[key + ".alphas"] = dec.alphas
            out[key + ".betas"] = dec.betas
            y = 0.1 * (np.arange(dec.steps_taken) + 1)
            out[key + ".x_y01"] = orc.lanczos_pass_two(kkt.a, b, dec, y)
            if flavour in ("nod", "wc"):  # moderate spectrum: exp is finite
                out[key + ".x_exp_k30"] = orc.lanczos_two_pass(kkt.a, b, 30, npo.exp_tk_solver)
            if flavour == "wc":  # inv inside the convergence plateau, b in range(A)
                bc = kkt.a.apply(np.full(n, 1.0 / np.sqrt(n)))
                out[key + ".b_range"] = bc
                out[key + ".x_inv_k200"] = orc.lanczos_two_pass(kkt.a, bc, 200, npo.inv_tk_solver)
    np.savez_compressed(os.path.join(HERE, "golden_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
