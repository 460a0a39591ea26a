#!/bin/bash
# per-warp-slot GN partials + smem-staged k_attn_mix: tests, reproducibility probe, per-launch profiles
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_d.log
timeout 200 python scripts/gpu_batch_dep.py 512 > $O/batch_dep_d.log 2>&1; echo "probe rc=$?"; grep "estimator.out\|first" $O/batch_dep_d.log
timeout 100 python scripts/gpu_profile_ops.py 32 512 tf32 > $O/ops_tf32_d.txt 2>&1; head -2 $O/ops_tf32_d.txt
timeout 100 python scripts/gpu_profile_ops.py 32 512 bf16 > $O/ops_bf16_d.txt 2>&1; head -2 $O/ops_bf16_d.txt
timeout 100 python scripts/gpu_profile_ops.py 1 512 tf32 > $O/ops_b1_tf32_d.txt 2>&1; head -2 $O/ops_b1_tf32_d.txt
timeout 100 python scripts/gpu_profile_ops.py 1 512 bf16 > $O/ops_b1_bf16_d.txt 2>&1; head -2 $O/ops_b1_bf16_d.txt
