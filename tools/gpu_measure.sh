#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/measure.log
: > $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_measure.json 2>> $LOG
cat gpurun_out/bench_measure.json >> $LOG
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py > gpurun_out/op_breakdown.txt 2>&1
head -30 gpurun_out/op_breakdown.txt >> $LOG
echo "=== ncu launch list (one eager step)" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py >> $LOG 2>&1
echo "=== splat bench" >> $LOG
timeout 600 python tools/splat_bench.py >> $LOG 2>&1
tail -n 60 $LOG
