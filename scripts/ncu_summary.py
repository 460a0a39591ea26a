"""Summarise an .ncu-rep (raw page) into a small markdown table: python scripts/ncu_summary.py rep.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
print(f"# ncu --set full summary: {rep}\n")
print("| # | kernel | " + " | ".join(w[1] for w in want) + " |")
print("|---|---|" + "---|" * len(want))
for n, r in enumerate(data):
    name = r[idx["Kernel Name"]].replace("void ", "").replace("sbk::", "").split("(sbk")[0].split("(Conv")[0]
    cells = []
    for k, _ in want:
        if k in idx:
            v = r[idx[k]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[idx[k]]}".strip())
        else:
            cells.append("-")
    print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
