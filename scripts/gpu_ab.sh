#!/bin/bash
# A/B of the remaining measurement knobs on one box, per precision mode: CTA pairs on / off (SBK_NO_PAIR) x accumulation-run
# length of the fp32x3 mode (SBK_X3_FLUSH, sub-stages per run).    usage: scripts/gpu_ab.sh "fp32x3 tf32"
set -u
O=gpurun_out; mkdir -p $O
MODES=${1:-"fp32x3 tf32"}
for m in $MODES; do
  for fl in 0 4; do
    SBK_X3_FLUSH=$fl SBK_NO_PAIR=1 timeout 120 python scripts/gpu_profile_ops.py 32 512 $m > $O/ab_${m}_fl${fl}_single.txt 2>&1; echo "flush=$fl single: $(head -1 $O/ab_${m}_fl${fl}_single.txt)"
    SBK_X3_FLUSH=$fl timeout 120 python scripts/gpu_profile_ops.py 32 512 $m > $O/ab_${m}_fl${fl}_pair.txt 2>&1; echo "flush=$fl pair:   $(head -1 $O/ab_${m}_fl${fl}_pair.txt)"
  done
done
