#!/bin/bash
# quick GPU check of a kernel change: the parity tests of the tensor-core paths (errors printed), then per-launch profiles
set -u
O=gpurun_out
TAG=${1:-quick}
MODES=${2:-"fp32x3 tf32 bf16"}
mkdir -p $O
timeout 900 python -m pytest tests/test_fp32x3_gpu.py tests/test_parity_gpu.py tests/test_hifigan.py tests/test_diffvc_gpu.py -m gpu -q -x -s > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest_$TAG.log
grep -E "^fp32x3 (kind|B=32)" $O/pytest_$TAG.log | head -20
for m in $MODES; do timeout 200 python scripts/gpu_profile_ops.py 32 512 $m > $O/ops_${m}_$TAG.txt 2>&1; head -n 2 $O/ops_${m}_$TAG.txt; done
