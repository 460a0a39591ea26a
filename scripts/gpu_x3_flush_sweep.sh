#!/bin/bash
# fp32x3 conv kernels: cost of the chunked accumulation (SBK_X3_FLUSH = sub-stages per accumulation run) + one ncu capture
mkdir -p gpurun_out
for f in 3 6 12 100000; do
  SBK_X3_FLUSH=$f timeout 200 python scripts/gpu_profile_ops.py 32 512 fp32x3 > gpurun_out/x3_ops_flush$f.txt 2>&1
  head -2 gpurun_out/x3_ops_flush$f.txt
  grep "downs.0.0.block2.raw\|downs.1.0.block2.raw\|downs.2.0.block2.raw\|downs.0.2.kvraw\|downs.0.3.out" gpurun_out/x3_ops_flush$f.txt
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_conv_tc_x3 --launch-skip 2 --launch-count 10 \
   -o gpurun_out/prof_x3_conv -f python scripts/gpu_one_call.py 32 512 fp32x3 > gpurun_out/ncu_x3.log 2>&1
tail -3 gpurun_out/ncu_x3.log
