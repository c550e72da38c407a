#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run30.log
: > $LOG
echo "=== normal" >> $LOG
timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj|geglu|ffout" >> $LOG
echo "=== no stores" >> $LOG
A3D_GEMM_NOSTORE=1 timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj|geglu|ffout" >> $LOG
echo "=== no stores trace qkv" >> $LOG
A3D_GEMM_NOSTORE=1 timeout 300 python tools/gemm_trace.py qkv 2>&1 | head -8 >> $LOG
tail -n 40 $LOG
