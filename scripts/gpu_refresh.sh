#!/bin/bash
# Round-end evidence refresh on one B200: parity tests, bench line, per-launch event profile, ncu launch list,
# ncu full capture of the tensor-core kernels, DRAM traffic of the 3x3 conv class.  Outputs -> gpurun_out/.
set -u
O=gpurun_out
TAG=${1:-r1f}
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" 
timeout 400 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
timeout 200 python scripts/gpu_profile_ops.py 32 512 tf32 > $O/ops_$TAG.txt 2>&1; echo "ops rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-fp32-leg > $O/bench_under_ncu_$TAG.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc|k_attn_kv' -s 0 -c 12 -o $O/prof_tc_$TAG -f \
    python scripts/gpu_profile_ops.py 32 512 tf32 > $O/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_conv_tc -s 0 -c 60 --csv --log-file $O/traffic_$TAG.csv python scripts/gpu_profile_ops.py 32 512 tf32 > $O/traffic_$TAG.log 2>&1; echo "traffic rc=$?"
tail -3 $O/pytest_gpu_$TAG.log; head -c 600 $O/bench_$TAG.json
