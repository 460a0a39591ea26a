#!/bin/bash
# Round-2 (second session) baseline: parity tests + per-launch profiles + full ncu captures of the fp32x3 kernels.
set -u
O=gpurun_out
TAG=${1:-r2b}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_gpu_$TAG.log
for m in fp32x3 tf32; do timeout 200 python scripts/gpu_profile_ops.py 32 512 $m > $O/ops_${m}_$TAG.txt 2>&1; head -n 2 $O/ops_${m}_$TAG.txt; done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc_x3|k_kv_ctx' -s 0 -c 24 -o $O/prof_x3_$TAG -f \
    python scripts/gpu_one_call.py 32 512 fp32x3 > $O/ncu_full_x3_$TAG.log 2>&1; echo "full x3 rc=$?"
