#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run26.log
: > $LOG
echo "=== ncu geglu l0" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 80 -c 1 -o gpurun_out/prof_geglu python tools/kernel_bench.py gemm >> $LOG 2>&1
echo "=== ncu attn5" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn5_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn5 python tools/kernel_bench.py attn0 >> $LOG 2>&1
tail -n 30 $LOG
