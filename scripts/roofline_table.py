"""Per-kernel-class roofline table from a per-launch profile (scripts/gpu_profile_ops.py output):
   python scripts/roofline_table.py profiles/r1_ops_tf32_v25.txt tf32 > profiles/r1_roofline_tf32.md
Algorithmic FLOPs / bytes per launch are the library's own accounting (sbk_profile_ops: each stage reads its inputs once and
writes its outputs once); peaks: HBM 6569.6 GB/s and cuBLAS bf16 1386.8 TFLOP/s sustained (MEASURED_PEAKS.json), tf32 = half."""
import collections
import json
import os
import re
import sys

path, prec = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = pk.get("hbm_gbs", 6569.6)
# tensor peak per ALGORITHMIC flop: bf16 = the cuBLAS bf16 rate, tf32 = half of it, fp32x3 = a quarter (two tensor-core passes
# at the tf32 instruction rate per MAC: one tf32 MMA + one fp16 correction MMA)
TEN = pk.get("bf16_tflops_sustained", 1386.8) * {"bf16": 1.0, "tf32": 0.5, "fp32x3": 0.25}[prec]


def klass(n):
    if n.endswith("block1.raw") and "downs.0.0." in n: return "first conv (CUDA cores, 2->64)"
    if n.endswith(".raw"): return "3x3 conv (tcgen05) + GN partials"
    if n.endswith(".act"): return "k_gn_act: GN + Mish + time bias -> operand"
    if "kvpart" in n: return "k_attn_kv(_x3): k/v projection + softmax partials + P V^T"
    if n.endswith(".ctx"): return "k_attn_ctx: merge partials"
    if n.endswith(".mix"): return "k_attn_mix: fold to_out ctx^T W_q"
    if re.search(r"\.2\.out|mid_attn.out", n): return "attention apply 1x1 (+ residual)"
    if re.search(r"downs\.\d\.3\.out", n): return "Downsample 3x3 s2"
    if re.search(r"ups\.\d\.3\.out", n): return "Upsample convT 4x4 s2"
    if n == "estimator.out": return "k_final: GN + Mish + 1x1 + Euler"
    return "ResnetBlock tail (k_resfinal / 1x1 res_conv epilogue)"


agg = collections.OrderedDict()
head = []
for line in open(path):
    if line.startswith("#"):
        head.append(line.strip("# \n"))
        continue
    f = line.split()
    name, ms, tf, gb = f[0], float(f[1]), float(f[3]), float(f[5])
    a = agg.setdefault(klass(name), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += tf * ms * 1e-3; a[3] += gb * ms * 1e-3      # TFLOP and GB totals
tot = sum(a[1] for a in agg.values())
print(f"# Roofline per kernel class, {prec}, B=32 T=512 ({'; '.join(head)})\n")
PEAK_NOTE = {"bf16": "cuBLAS bf16 sustained", "tf32": "half the cuBLAS bf16 sustained rate",
             "fp32x3": "a quarter of the cuBLAS bf16 sustained rate: two tf32-rate passes per MAC"}[prec]
print(f"Peaks: HBM {HBM:.0f} GB/s, tensor {TEN:.0f} TFLOP/s ({PEAK_NOTE}, MEASURED_PEAKS.json).")
print("`frac` = achieved / peak of the BINDING resource (the larger of the two fractions).\n")
print("| class | launches | ms | share | TFLOP/s | GB/s (algorithmic) | tensor frac | HBM frac | bound |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
for k, (n, ms, tfl, gbt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf, gb = tfl / (ms * 1e-3), gbt / (ms * 1e-3)
    ft, fh = tf / TEN, gb / HBM
    bound = "tensor" if ft >= fh else "HBM"
    if max(ft, fh) < 0.25:
        bound = "latency / issue"
    print(f"| {k} | {n} | {ms:.3f} | {ms / tot:.3f} | {tf:.0f} | {gb:.0f} | {ft:.2f} | {fh:.2f} | {bound} |")
print(f"| **all** | {sum(a[0] for a in agg.values())} | {tot:.3f} | 1.000 | {sum(a[2] for a in agg.values()) / (tot * 1e-3):.0f} | {sum(a[3] for a in agg.values()) / (tot * 1e-3):.0f} | | | |")
