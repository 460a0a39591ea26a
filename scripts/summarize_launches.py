"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import io
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for x in csv.DictReader(io.StringIO("".join(rows))):
    k = x["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(x["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(x["Metric Unit"], 1.0)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list summary: {sys.argv[1]}")
print("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
print(f"| **all** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 1.000 |")
