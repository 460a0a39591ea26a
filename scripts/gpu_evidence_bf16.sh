#!/bin/bash
# bf16 evidence run: parity tests (fp32 + tf32 + bf16, Grad-TTS + DiffVC), per-launch bf16 profile, config 3, DiffVC config 4
# in bf16, then the headline bench line.  Outputs -> gpurun_out/.
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu_b.log
timeout 200 python scripts/gpu_profile_ops.py 32 512 bf16 > $O/ops_bf16_b.txt 2>&1; echo "ops rc=$?"; head -2 $O/ops_bf16_b.txt
timeout 300 python scripts/gpu_config3.py > $O/config3_bf16.jsonl 2> $O/config3.err; echo "config3 rc=$?"; cat $O/config3_bf16.jsonl
timeout 300 python scripts/gpu_diffvc_bench.py bf16 > $O/diffvc_config4_bf16.txt 2>&1; echo "diffvc rc=$?"; head -2 $O/diffvc_config4_bf16.txt
timeout 400 python bench.py > $O/bench_b.json 2> $O/bench_b.err; echo "bench rc=$?"; head -c 300 $O/bench_b.json
