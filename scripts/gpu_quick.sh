#!/bin/bash
# quick GPU check of a kernel change: the parity tests of the tensor-core paths, then per-launch profiles of the three modes
set -u
O=gpurun_out
TAG=${1:-quick}
mkdir -p $O
timeout 900 python -m pytest tests/test_fp32x3_gpu.py tests/test_parity_gpu.py tests/test_hifigan.py tests/test_diffvc_gpu.py -m gpu -q -x > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_$TAG.log
for m in fp32x3 tf32 bf16; do timeout 200 python scripts/gpu_profile_ops.py 32 512 $m > $O/ops_${m}_$TAG.txt 2>&1; head -n 2 $O/ops_${m}_$TAG.txt; done
