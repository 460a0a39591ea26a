#!/bin/bash
# reproducibility fix check + bf16 evidence: tests, batch-dependence probe, B=1 latency profiles, sweeps, ncu captures (bf16)
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_c.log
timeout 200 python scripts/gpu_batch_dep.py 512 > $O/batch_dep_after.log 2>&1; echo "probe rc=$?"; grep "estimator.out\|first" $O/batch_dep_after.log
timeout 100 python scripts/gpu_profile_ops.py 1 512 tf32 > $O/ops_b1_tf32.txt 2>&1; head -2 $O/ops_b1_tf32.txt
timeout 100 python scripts/gpu_profile_ops.py 1 512 bf16 > $O/ops_b1_bf16.txt 2>&1; head -2 $O/ops_b1_bf16.txt
timeout 100 python scripts/gpu_profile_ops.py 32 512 tf32 > $O/ops_tf32_c.txt 2>&1; head -2 $O/ops_tf32_c.txt
timeout 300 python scripts/gpu_sweep.py tf32 > $O/sweep_tf32.jsonl 2>&1; echo "sweep tf32 rc=$?"
timeout 300 python scripts/gpu_sweep.py bf16 > $O/sweep_bf16.jsonl 2>&1; echo "sweep bf16 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc|k_attn_kv|k_gn_act_bf16|k_resfinal_bf16' -s 0 -c 14 -o $O/prof_bf16 -f \
    python scripts/gpu_profile_ops.py 32 512 bf16 > $O/ncu_full_bf16.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bf16.csv \
    python bench.py --precision bf16 --steps 2 --warmup 1 --no-fp32-leg > $O/bench_under_ncu_bf16.log 2>&1; echo "launches rc=$?"
