#!/bin/bash
# full ncu capture of the kernels matching a regex during one estimator call:
#   scripts/gpu_ncu_kernel.sh REGEX TAG [COUNT] [precision] [SKIP]      (environment variables such as SBK_NO_PAIR pass through)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${5:-0} -c ${3:-8} -o $O/prof_$2 -f \
    python scripts/gpu_one_call.py 32 512 ${4:-fp32x3} > $O/ncu_$2.log 2>&1; echo "ncu rc=$?"; tail -n 2 $O/ncu_$2.log
