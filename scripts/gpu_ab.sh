#!/bin/bash
# A/B of the measurement knobs on one box: L2 prefetch distance (SBK_PFK) x CTA pairs (SBK_NO_PAIR), per precision mode
set -u
O=gpurun_out; mkdir -p $O
MODES=${1:-"fp32x3 tf32"}
for m in $MODES; do
  for pf in 0 4; do
    SBK_PFK=$pf SBK_NO_PAIR=1 timeout 120 python scripts/gpu_profile_ops.py 32 512 $m > $O/ab_${m}_pf${pf}_single.txt 2>&1; echo "pf=$pf single: $(head -1 $O/ab_${m}_pf${pf}_single.txt)"
    SBK_PFK=$pf timeout 120 python scripts/gpu_profile_ops.py 32 512 $m > $O/ab_${m}_pf${pf}_pair.txt 2>&1; echo "pf=$pf pair:   $(head -1 $O/ab_${m}_pf${pf}_pair.txt)"
  done
done
