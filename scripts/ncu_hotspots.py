"""Top SASS instructions by warp-stall samples for launch N of an .ncu-rep source page dump:
   ncu -i rep --page source --csv > src.csv ; python scripts/ncu_hotspots.py src.csv N [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.2
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[n]
hdr, data = b["rows"][0], b["rows"][1:]
isrc, iall, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[iall] or 0) for r in data)
print(b["name"], "total samples", tot)
for i, r in enumerate(data):
    s = int(r[iall] or 0)
    if s > tot * thr / 100:
        print(i, r[isrc][:90], s, f"{100 * s / tot:.1f}%", r[iex])
