#!/bin/bash
# Round-2 evidence refresh on one B200 (outputs -> gpurun_out/; the summaries that matter are copied to profiles/):
# parity tests, the bench line of both arms, per-launch event profiles of the three tensor-core modes, the ncu launch list of
# the bench command, full ncu captures of the dominant kernels, DRAM traffic of the 3x3 conv class of the benched binary.
set -u
O=gpurun_out
TAG=${1:-r2}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest_gpu_$TAG.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "bench reference rc=$?"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
for m in fp32x3 tf32 bf16; do timeout 200 python scripts/gpu_profile_ops.py 32 512 $m > $O/ops_${m}_$TAG.txt 2>&1; head -n 2 $O/ops_${m}_$TAG.txt; done
timeout 200 python scripts/gpu_profile_ops.py 1 512 fp32x3 > $O/ops_fp32x3_b1_$TAG.txt 2>&1; head -n 1 $O/ops_fp32x3_b1_$TAG.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-extra-legs > $O/bench_under_ncu_$TAG.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc_x3|k_attn_kv_x3' -s 0 -c 16 -o $O/prof_x3_$TAG -f \
    python scripts/gpu_one_call.py 32 512 fp32x3 > $O/ncu_full_x3_$TAG.log 2>&1; echo "full x3 rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_conv_tc_x3' -s 0 -c 60 --csv --log-file $O/traffic_x3_$TAG.csv python scripts/gpu_one_call.py 32 512 fp32x3 > $O/traffic_x3_$TAG.log 2>&1; echo "traffic rc=$?"
head -c 600 $O/bench_$TAG.json; echo; head -c 400 $O/bench_ref_$TAG.json
