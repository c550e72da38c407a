#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run2.log
: > $LOG
echo "=== smoke" >> $LOG
timeout 600 python __graft_entry__.py --smoke >> $LOG 2>&1
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | tail -n 40 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2>> $LOG
cat gpurun_out/bench_r01.json >> $LOG
echo "=== ncu launch list (eager, 1 step)" >> $LOG
A3D_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
   python tools/one_step.py >> $LOG 2>&1
tail -n 5 gpurun_out/launches_r01.csv >> $LOG
tail -n 80 $LOG
