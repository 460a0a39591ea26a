"""DRAM traffic per launch of the dominant kernel class from an ncu CSV (metrics dram__bytes_read.sum,
dram__bytes_write.sum, gpu__time_duration.sum): writes profiles/<name>.json for bench.py's roofline.traffic.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:k_conv_tc -s <skip> -c <n> --csv --log-file gpurun_out/traffic.csv python scripts/gpu_profile_ops.py 32 512 tf32
    python scripts/ncu_traffic.py gpurun_out/traffic.csv "k_conv_tc<1" profiles/r1_traffic_conv3x3.json
"""
import collections
import csv
import io
import json
import sys

path, pattern, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [l for l in open(path) if l.startswith('"')]
per = collections.defaultdict(dict)
for x in csv.DictReader(io.StringIO("".join(rows))):
    if pattern not in x["Kernel Name"]:
        continue
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
    per[x["ID"]][x["Metric Name"]] = v
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
tm = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import csrc_digest   # ties the capture to the kernel sources it was taken from (bench.py reports it only if they match)
res = {"csrc_digest": csrc_digest(), "source": path, "kernel_pattern": pattern, "launches": n, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
       "dram_read_per_launch": rd / max(n, 1), "dram_write_per_launch": wr / max(n, 1), "ncu_seconds_per_launch": tm / max(n, 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
