"""Top stall sites of an `ncu --page source --csv` export.  usage: ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
iS = hdr.index("# Samples"); isrc = hdr.index("Source")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try: return int(x)
    except Exception: return 0
tot = sum(num(r[iS]) for r in data)
agg = {s: sum(num(r[hdr.index(s)]) for r in data) for s in stalls}
print("total samples", tot)
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -num(r[iS]))[:N]:
    st = sorted(((s, num(r[hdr.index(s)])) for s in stalls), key=lambda x: -x[1])[:2]
    print(f"{num(r[iS]):6d} {100.0*num(r[iS])/tot:5.1f}% {r[isrc][:105]:105s} {st}")
