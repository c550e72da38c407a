#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run13.log
: > $LOG
echo "=== op breakdown" >> $LOG
timeout 600 python tools/op_breakdown.py >> $LOG 2>&1
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01f.json 2>> $LOG
cat gpurun_out/bench_r01f.json >> $LOG
tail -n 100 $LOG
