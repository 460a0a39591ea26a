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
# libsbk's own kernels only (bench.py also times a cuBLAS TF32 matmul - its peak probe - and torch draws the inputs), and the
# kernels of the reverse step proper (without the once-per-call set-up: time table, xt = z*mask, graph bookkeeping)
own = {k: v for k, v in agg.items() if k.startswith("sbk::") or k.startswith("<unnamed>::k_set")}
setup = ("sbk::k_time_table", "sbk::k_scale_mask", "sbk::k_spk", "<unnamed>::k_set")
step = {k: v for k, v in own.items() if not k.startswith(setup)}
ts = sum(v[1] for v in step.values())
print("\n## kernels of the reverse step only (libsbk, once-per-call set-up excluded)\n")
print("| kernel | launches | total us | share of the step |\n|---|---:|---:|---:|")
for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / ts:.3f} |")
c3 = sum(v[1] for k, v in step.items() if "_pair<" in k or "k_conv_tc_x3<1," in k or "k_conv_tc<1," in k or "k_first_conv" in k)
print(f"| **3x3 conv class (incl. k_first_conv)** | | {c3:.1f} | {c3 / ts:.3f} |")
