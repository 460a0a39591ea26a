"""Opcode histogram of libsbk.so's SASS, per kernel (evidence that the hot kernels are tcgen05 / TMEM / bulk-copy code).

    python scripts/sass_histogram.py > profiles/r2_sass_opcodes.md        # CPU only: cuobjdump reads the in-tree .so

Counts the Blackwell-specific mnemonics the profiling recipe names: UTCHMMA / UTCQMMA (tcgen05.mma), UTCBAR (tcgen05.commit),
LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTMALDG (tensor-map TMA loads), LDGSTS (cp.async), SYNCS (mbarrier),
plus FFMA / MUFU for the CUDA-core kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-backbones_b200", "libsbk.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA", "MUFU", "HFMA2", "LDS", "STS", "LDG", "STG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = demangle(list(counts))
    print("# SASS opcode histogram of libsbk.so (sm_100a), per kernel\n")
    print("`python scripts/sass_histogram.py` (cuobjdump -sass on the in-tree library). Only kernels with at least one instruction are listed;")
    print("template arguments are `<GEOM, BF16, NT, RES>` for `k_conv_tc` (GEOM 0 = 1x1, 1 = 3x3, 2 = Downsample, 3 = Upsample, 4/5/6 = Conv1d K=3/7/11)")
    print("and `<GEOM, NT, RES>` for the fp32x3 variants `k_conv_tc_x3`.\n")
    print("| kernel | instrs | " + " | ".join(OPS) + " |")
    print("|---|---|" + "---|" * len(OPS))
    tot = collections.Counter()
    for k, c in counts.items():
        if c["_total"] == 0:
            continue
        short = re.sub(r"\(.*\)$", "", names.get(k, k)).replace("sbk::", "").replace("void ", "")
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
        tot.update(c)
    print(f"| **all kernels** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
    print("\nNo `UTMALDG`: operand tiles are moved with 1-D `cp.async.bulk` (UBLKCP) runs, which is what the channel-chunk-planar")
    print("HBM layout makes possible (a conv halo row of one 16-byte channel chunk is one contiguous run); no tensor maps are needed.")


if __name__ == "__main__":
    main()
