"""Seeded "netgen-shaped" KKT instance generator (replaces the reference's `datagen` binary).

The reference shells out to pargen -> netgen -> qfcgen (src/bin/datagen.rs:109-126).  Those tools
cannot be used for the benchmark inputs: netgen has static limits of 100k nodes / 1.1M arcs
(data/netgen/src/netgen.h:76-77), pargen seeds from time(NULL) (data/qcnd/pargen.c:52-55) so instances
are not reproducible, and qfcgen's 3-line `.qfc` makes the reference loader drop D.
This module reproduces the *shape* of their output:

  * p = floor((1 + sqrt(1 + 8m/rho'))/2), rho' in {.25,.5,.75}             (pargen.c:41-50)
  * #sources, #sinks ~ randint[1, 0.1p]; sinks are the highest-numbered nodes and have no out-arcs,
    sources (lowest-numbered) have no in-arcs                               (pargen.c:73-77, netgen)
  * arcs are emitted grouped by tail, tails increasing, heads in random order, no self-loops, no
    duplicate (tail, head) pairs                                            (netgen.c:305-322,432-438)
  * quadratic cost ("aa"): D = 1 + U[100c, 1000c], c = 1 + randint[3b, 10b), b = arc cost
    ~ randint[1, maxcost], maxcost ~ randint[10, 108]   (qfcgen.c:184-196, pargen.c:83-85)
    or ("wc"): D ~ U[1, 10]                               (tex/report.tex:338-342)

and writes `.dmx` + `.qfc` with ONE VALUE PER LINE (2m+1 lines) so that the reference loader's
line semantics (src/utils/data_loader.rs:166-198) really populate D.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

_RHO = {1: 0.25, 2: 0.5, 3: 0.75}


@dataclass
class KKTInstance:
    m: int            # arcs
    p: int            # nodes
    tail: np.ndarray  # uint32[m], 0-based node index
    head: np.ndarray  # uint32[m]
    d: np.ndarray     # float64[m] quadratic costs (diagonal block D)
    fixed: np.ndarray  # float64[m] fixed costs (ignored by the loader; lines 1..m of the .qfc)
    rho: int
    seed: int
    flavour: str

    @property
    def n(self) -> int:
        return self.m + self.p

    @property
    def name(self) -> str:  # src/bin/datagen.rs:109-117
        return f"netgen-{self.m}-{self.rho}-{self.seed}-a-a-ns"


def num_nodes(m: int, rho: int = 3) -> int:
    return int(math.floor((1.0 + math.sqrt(1.0 + (8.0 * m) / _RHO[rho])) / 2.0))


def gen_kkt(m: int, rho: int = 3, seed: int = 1, flavour: str = "aa") -> KKTInstance:
    rng = np.random.Generator(np.random.PCG64(seed))
    p = num_nodes(m, rho)
    max_nodes = max(1, int(0.1 * p))
    nsrc = int(rng.integers(1, max_nodes + 1))
    nsink = int(rng.integers(1, max_nodes + 1))
    ntail = p - nsink
    if ntail < 1:
        raise ValueError("instance too small")
    # capacity of every tail: candidate heads are the non-source nodes except itself
    tails = np.arange(ntail)
    cap = (p - nsrc) - (tails >= nsrc).astype(np.int64)
    if m < ntail or m > int(cap.sum()):
        raise ValueError(f"cannot place {m} arcs on {p} nodes")
    # out-degrees: one skeleton arc per tail + a uniform multinomial split of the rest, clipped to capacity
    deg = 1 + rng.multinomial(m - ntail, np.full(ntail, 1.0 / ntail))
    while True:
        over = deg - cap
        excess = int(over[over > 0].sum())
        if excess == 0:
            break
        deg = np.minimum(deg, cap)
        room = cap - deg
        deg = deg + rng.multivariate_hypergeometric(room, excess)
    assert int(deg.sum()) == m
    tail = np.repeat(tails, deg).astype(np.uint32)
    head = np.empty(m, dtype=np.uint32)
    cand_all = np.arange(nsrc, p, dtype=np.uint32)
    pos = 0
    for t in range(ntail):
        k = int(deg[t])
        cand = cand_all if t < nsrc else np.delete(cand_all, t - nsrc)
        head[pos:pos + k] = rng.choice(cand, size=k, replace=False)
        pos += k
    if flavour == "aa":
        maxcost = int(rng.integers(10, 109))
        b = rng.integers(1, maxcost + 1, size=m)
        cc = 1.0 + (3 * b + rng.integers(0, 7 * b))          # Cc = rand % (10b-3b) + 3b + 1
        d = 1.0 + 100.0 * cc + rng.random(m) * (900.0 * cc)   # Ca = U[100Cc, 1000Cc] + 1
        fixed = cc.astype(np.float64)
    elif flavour == "wc":
        d = 1.0 + 9.0 * rng.random(m)
        fixed = np.ones(m)
    else:
        raise ValueError("flavour must be 'aa' or 'wc'")
    return KKTInstance(m, p, tail, head, d.astype(np.float64), fixed, rho, seed, flavour)


def write_dmx(path: str, inst: KKTInstance) -> None:
    """DIMACS min-cost-flow layout as emitted by netgen (data/netgen/src/netgen.c:494-557)."""
    lines = [
        "c NETGEN-shaped flow network (seeded generator, two_pass_lanczos_b200.datagen)",
        f"c  seed {inst.seed}  rho {inst.rho}  flavour {inst.flavour}",
        f"p min {inst.p} {inst.m}",
        f"n 1 {100}",
        f"n {inst.p} {-100}",
    ]
    t1 = inst.tail.astype(np.int64) + 1
    h1 = inst.head.astype(np.int64) + 1
    body = "\n".join(f"a {t} {h} 0 100 1" for t, h in zip(t1.tolist(), h1.tolist()))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n" + body + "\n")


def write_qfc(path: str, inst: KKTInstance, layout: str = "lines", d_len: int | None = None) -> None:
    """`.qfc`: "lines" = loader-compatible 2m+1-line layout; "qfcgen" = the 3-line layout qfcgen
    really writes (data/qcnd/qfcgen.c:210-218), which makes the reference loader return an empty D."""
    d = inst.d if d_len is None else inst.d[:d_len]
    with open(path, "w") as f:
        f.write(f"{inst.m}\n")
        if layout == "lines":
            f.write("\n".join(repr(float(v)) for v in inst.fixed) + "\n")
            if len(d):
                f.write("\n".join(repr(float(v)) for v in d) + "\n")
        elif layout == "qfcgen":
            f.write(" ".join(f"{v:f}" for v in inst.fixed) + " \n")
            f.write(" ".join(f"{v:f}" for v in d) + " \n")
        else:
            raise ValueError(layout)


def write_instance(directory: str, inst: KKTInstance, layout: str = "lines"):
    os.makedirs(directory, exist_ok=True)
    dmx = os.path.join(directory, inst.name + ".dmx")
    qfc = os.path.join(directory, inst.name + ".qfc")
    write_dmx(dmx, inst)
    write_qfc(qfc, inst, layout)
    return dmx, qfc


def kkt_csc(inst: KKTInstance):
    """Host CSC (colptr u64, rowidx u64, val f64) of A=[[D,E^T],[E,0]] exactly as the reference loader
    builds it (src/utils/data_loader.rs:222-251): arcs first, nodes last, rows ascending per column."""
    import scipy.sparse as sp

    m, p = inst.m, inst.p
    j = np.arange(m)
    t = inst.tail.astype(np.int64)
    h = inst.head.astype(np.int64)
    rows = np.concatenate([j, m + t, m + h, j, j])
    cols = np.concatenate([j, j, j, m + t, m + h])
    ones = np.ones(m)
    vals = np.concatenate([inst.d, ones, -ones, ones, -ones])
    a = sp.csc_matrix((vals, (rows, cols)), shape=(m + p, m + p))
    a.sort_indices()
    return a.indptr.astype(np.uint64), a.indices.astype(np.uint64), a.data.astype(np.float64)
