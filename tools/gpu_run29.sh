#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run29.log
: > $LOG
for w in qkv geglu sqkv proj; do
echo "=== gemm trace $w" >> $LOG
timeout 300 python tools/gemm_trace.py $w >> $LOG 2>&1
done
tail -n 90 $LOG
