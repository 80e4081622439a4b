"""Stall samples of an `ncu --page source --csv` export grouped by code region (regions end at BAR.SYNC), per kernel.
usage: ncu_regions.py file.csv [top-per-region]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kernels, hdr, cur = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if cur is not None and hdr and len(r) == len(hdr):
        cur["rows"].append(r)
seen = set()
iS, isrc = hdr.index("# Samples"), hdr.index("Source")
iex = hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sidx = {s: hdr.index(s) for s in stalls}


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


for k in kernels:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    data = k["rows"]
    tot = sum(num(r[iS]) for r in data)
    print(f"=== {k['name'][:60]}  instructions {len(data)}  samples {tot}")
    region, start = [], 0
    regions = []
    for i, r in enumerate(data):
        region.append(r)
        if "BAR.SYNC" in r[isrc] or i == len(data) - 1:
            regions.append((start, i, region))
            region, start = [], i + 1
    for a, b, reg in regions:
        s = sum(num(r[iS]) for r in reg)
        if s < tot * 0.005:
            continue
        agg = sorted(((st, sum(num(r[sidx[st]]) for r in reg)) for st in stalls), key=lambda x: -x[1])[:4]
        ex = sum(num(r[iex]) for r in reg)
        print(f"-- instr {a}..{b}: {s} samples ({100.0 * s / tot:.1f}%), warp-instr executed {ex}, " +
              ", ".join(f"{n[6:]} {v}" for n, v in agg))
        for r in sorted(reg, key=lambda r: -num(r[iS]))[:N]:
            st = sorted(((x, num(r[sidx[x]])) for x in stalls), key=lambda x: -x[1])[:2]
            print(f"      {num(r[iS]):6d}  {r[isrc].strip()[:70]:70s} {st[0][0][6:]} {st[0][1]} {st[1][0][6:]} {st[1][1]}")
