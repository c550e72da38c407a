#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run17.log
: > $LOG
echo "=== ncu attention v4 d=40" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn4 python tools/kernel_bench.py attn0 >> $LOG 2>&1
tail -n 20 $LOG
