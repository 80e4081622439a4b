#!/usr/bin/env python
"""Text pair (.dmx + .qfc) -> TPLKKT1 container, or a generated instance straight to a container.

    python scripts/kkt_convert.py net.dmx net.qfc net.tplkkt
    python scripts/kkt_convert.py --gen 50000000 --rho 3 --flavour aa out.tplkkt

Prints the load time of both forms (the container is read at file speed)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("paths", nargs="+")
    ap.add_argument("--gen", type=int, default=0, help="arcs of a netgen-shaped instance to generate instead of reading text")
    ap.add_argument("--rho", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--flavour", default="aa")
    args = ap.parse_args()
    from two_pass_lanczos_b200 import data_loader, datagen

    if args.gen:
        (out,) = args.paths
        inst = datagen.gen_kkt(args.gen, args.rho, args.seed, args.flavour)
        t = time.time()
        data_loader.write_kkt_binary(out, inst.p, inst.tail, inst.head, inst.d)
        print(f"generated {inst.m} arcs / {inst.p} nodes -> {out} ({os.path.getsize(out) >> 20} MiB) in {time.time() - t:.2f} s")
    else:
        dmx, qfc, out = args.paths
        t = time.time()
        host = data_loader.load_kkt_host(dmx, qfc)
        t_text = time.time() - t
        host.save_binary(out)
        print(f"text pair: {host.num_arcs} arcs / {host.num_nodes} nodes, {host.num_costs} costs, loaded in {t_text:.2f} s")
    t = time.time()
    back = data_loader.load_kkt_host_binary(out)
    print(f"container: {back.num_arcs} arcs loaded in {time.time() - t:.3f} s")


if __name__ == "__main__":
    main()
